def _needs_real_shapely(*a, **k):
    raise ImportError("shapely is not installed: generating the Fourier-shape data sets (data.py sample_joint) needs the real package; "
                      "run tools/make_synthetic_data.py to create synthetic data/*.npy files of the right shapes instead")


Point = box = Polygon = LineString = MultiPolygon = _needs_real_shapely
