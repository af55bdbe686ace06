def rgb2hex(c, keep_alpha=False):
    r, g, b = (int(round(255 * float(v))) for v in tuple(c)[:3])
    return f"#{r:02x}{g:02x}{b:02x}"


def to_rgba(c, alpha=None):
    return tuple(c) if not isinstance(c, str) else (0.0, 0.0, 0.0, 1.0)
