"""Minimal stand-in for matplotlib (absent from this image) so that the reference's
simulation/models.py can be imported by oracle/make_golden.py.  Test infrastructure only."""
def use(*a, **k):
    pass
