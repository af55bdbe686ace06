"""Stand-in for shapely (reference: data.py:11-12).  The geometry is only needed to GENERATE the lens / plus Fourier-shape data
sets (data.py:88-94, 205-206); with data/<name>_{x,y}_{train,test}.npy present (tools/make_synthetic_data.py) it is never called."""
from . import geometry, ops  # noqa: F401
