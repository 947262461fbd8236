class Box:
    def __init__(self, *a, **k):
        pass


class Dict(dict):
    pass
