"""Stand-in for the `visdom` package (reference: monitoring.py:4,56-133,140-145): accepts every call the reference's
LiveVisualizer makes and sends nothing.  `check_connection()` answers True because the reference's non-live `Visualizer`
lacks the methods its own training loops call (`update_progress`, 3-argument `update_losses`): only the live one runs."""


class Visdom:
    def __init__(self, *args, **kwargs):
        self.env = kwargs.get("env")

    def check_connection(self, *args, **kwargs):
        return True

    def close(self, *args, **kwargs):
        return None

    def _send(self, *args, **kwargs):
        return None

    def matplot(self, *args, **kwargs):
        return "matplot_window"

    def scatter(self, *args, **kwargs):
        return "scatter_window"

    def __getattr__(self, name):            # any other plotting call: accepted, ignored
        return lambda *a, **k: None
