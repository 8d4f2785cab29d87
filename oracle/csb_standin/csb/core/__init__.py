from collections import OrderedDict  # noqa: F401


def iterable(obj):
    try:
        iter(obj)
        return True
    except TypeError:
        return False
