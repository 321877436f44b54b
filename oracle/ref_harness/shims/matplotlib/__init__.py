"""Import shim for matplotlib (absent from this image); plotting is out of scope."""
