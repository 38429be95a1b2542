"""Stand-in for sydr/io/visualisation.py (bokeh / panel HTML report, out of scope): run() does nothing."""


class Visualisation:
    def __init__(self, configuration):
        self.configuration = configuration

    def run(self):
        return None
