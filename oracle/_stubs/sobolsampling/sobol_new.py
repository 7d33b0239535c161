"""Stand-in for the (absent) `sobolsampling` package: start points only."""
from scipy.stats import qmc


def sobol_points(n, d):
    return qmc.Sobol(d, scramble=False).random(n)
