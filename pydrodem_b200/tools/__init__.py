"""Mirrors of the reference's `tools/` functions that sit on the hot path (same names, argument
meaning and file naming), running on libhydrocore instead of WhiteboxTools/scipy."""
