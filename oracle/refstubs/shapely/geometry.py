class Polygon:
    def __init__(self, *a, **k):
        pass


class Point:
    def __init__(self, *a, **k):
        pass
