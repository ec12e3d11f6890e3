"""TEST-ONLY stand-in: the reference drivers import matplotlib.pyplot at module level (main.py:27) and never call it on
the paths the tests touch; the image has no matplotlib."""
