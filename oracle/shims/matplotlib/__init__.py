"""Inert stand-in: the reference imports matplotlib at module top (params.py:5) but the hot
path never plots.  Test infrastructure only."""
