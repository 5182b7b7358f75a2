"""Import-time stand-in for matplotlib (reference: models/utils.py:3-5, utils/common_utils.py:4,8)."""


def use(*a, **k):
    pass
