"""Stand-in for sydr/enlightengui.py (terminal progress display, out of scope): the calls main.py and the
receiver make are accepted and ignored."""


class EnlightenGUI:
    def __init__(self):
        self.stage, self.status = None, None

    def updateMainStatus(self, stage: str, status: str):
        self.stage, self.status = stage, status

    def createReceiverGUI(self, receiver):
        pass

    def updateReceiverGUI(self, receiver):
        pass

    def stop(self):
        pass
