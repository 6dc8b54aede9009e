__version__ = "0.1.0"
ABI_VERSION = 1
