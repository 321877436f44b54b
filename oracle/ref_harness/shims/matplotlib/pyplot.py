def __getattr__(name):
    raise RuntimeError("matplotlib is not available in this image (plot=True is out of scope)")
