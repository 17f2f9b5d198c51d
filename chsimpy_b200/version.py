__version__ = "0.1.0"
# API level of the reference this package is a drop-in for (chsimpy/version.py)
__reference_version__ = "1.4.3"
