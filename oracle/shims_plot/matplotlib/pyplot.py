"""TEST-ONLY empty stand-in for matplotlib.pyplot (see __init__.py)."""
