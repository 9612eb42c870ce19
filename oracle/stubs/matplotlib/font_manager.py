class FontProperties:
    def __init__(self, *args, **kwargs):
        pass

    def set_size(self, *args, **kwargs):
        pass
