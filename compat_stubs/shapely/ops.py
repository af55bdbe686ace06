from .geometry import _needs_real_shapely

nearest_points = unary_union = _needs_real_shapely
