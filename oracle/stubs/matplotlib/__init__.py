"""Import-time stand-in for matplotlib (absent in this image). Plotting is out of scope."""
rcParams = {}


def use(*args, **kwargs):
    return None
