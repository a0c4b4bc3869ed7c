"""Minimal stand-in for the `accelerate` package (absent from this image, no network).

TEST INFRASTRUCTURE ONLY: lets oracle/make_golden.py import the unmodified reference from
/root/reference in THIS container so that golden vectors can be generated from it.  It provides only
the handful of helpers kronfluence imports; none of it is used by kronfluence_b200.
"""

# transformers probes `accelerate.__version__` when it finds the package on the path: a parseable, too-old version makes
# it treat accelerate as absent instead of failing on "N/A".
__version__ = "0.0.0"

