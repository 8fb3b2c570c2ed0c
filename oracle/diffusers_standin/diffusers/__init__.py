"""Minimal stand-in for the handful of `diffusers==0.24.0` symbols the reference's hot-path modules
import (requirements.txt:7 of the reference; the real package is not installed in this image and
there is no network).  TEST INFRASTRUCTURE ONLY: it exists so that `oracle/make_golden.py` and the
oracle cross-check tests can import the reference's own modules from /root/reference unmodified.
The behaviour of each class restates the 0.24.0 release from memory (see oracle/README.md);
nothing in the product path imports this package.
"""
__version__ = "0.24.0-standin"
