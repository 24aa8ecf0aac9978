class PyTorchLightningPruningCallback:
    def __init__(self, *a, **k):
        pass
