class Line2D:  # engine.py:5 imports the name only
    pass
