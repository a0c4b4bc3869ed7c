"""Minimal stand-in for the `accelerate` package (absent from this image, no network).

TEST INFRASTRUCTURE ONLY: lets oracle/make_golden.py import the unmodified reference from
/root/reference in THIS container so that golden vectors can be generated from it.  It provides only
the handful of helpers kronfluence imports; none of it is used by kronfluence_b200.
"""
