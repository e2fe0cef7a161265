"""Mirrors of the reference's `models/` entry points that sit on the hot path (array / raster level only)."""
