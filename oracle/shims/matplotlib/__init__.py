"""Stand-in for matplotlib (missing here): the validators import pyplot only to close figures and plot."""
