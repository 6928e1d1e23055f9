def __getattr__(name):
    raise RuntimeError("shim matplotlib.pyplot")
