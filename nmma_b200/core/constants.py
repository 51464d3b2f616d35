"""Constants of the hot path (subset of ``nmma/core/constants.py``)."""
import numpy as np

c_SI = 299792458.0
c_cgs = c_SI * 100.0
SENTINEL = float(np.nan_to_num(-np.inf))  # -1.7976931348623157e308, nmma/core/base.py:82
