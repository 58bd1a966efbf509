"""Inert stand-in for shapely (only imported by xopto.pf.util maps)."""
