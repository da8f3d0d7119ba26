"""Data preparation for the short-GRB sample: the reference's
``code/clean_data.py`` and ``code/kcorr.py`` as importable functions."""
from .clean_data import clean_raw, sgrbs, write_clean_csv          # noqa: F401
from .kcorr import WMAP9, k_correction, k_correct_grb, luminosity_distance_cm   # noqa: F401
