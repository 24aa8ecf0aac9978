class Accuracy:
    def __init__(self, *a, **k):
        pass
