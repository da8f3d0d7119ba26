"""``code/clean_data.py`` as functions: strip the plot-package rows from a Swift ``*_raw.txt`` light
curve and name the six columns."""
import os

import numpy as np

sgrbs = ["050724", "051016B", "051227", "060614", "061006", "061210", "070714B",      # clean_data.py:6-8
         "071227", "080123", "080503", "100212A", "100522A", "111121A",
         "150424A", "160410A"]

COLUMNS = ("t", "tpos", "tneg", "flux", "fluxpos", "fluxneg")


def clean_raw(infile):
    """clean_data.py:24-30: ``np.loadtxt(infile, comments=["!", "NO", "READ"])`` -> dict of six columns."""
    data = np.loadtxt(infile, comments=["!", "NO", "READ"], ndmin=2)
    return {name: data[:, i] for i, name in enumerate(COLUMNS)}


def write_clean_csv(outfile, cols):
    """clean_data.py:33: ``DataFrame.to_csv(outfile, index=False)`` with the column order above."""
    import pandas as pd
    os.makedirs(os.path.dirname(os.path.abspath(outfile)), exist_ok=True)
    pd.DataFrame(data={k: cols[k] for k in COLUMNS}).to_csv(outfile, index=False)
