class Callback:
    pass
