"""Verbosity-gated printing (host utility; mirrors reference pomdpy/util/console.py:12-28)."""
VERBOSITY = 3      # 0 fatal, 1 critical, 2 info, 3 loud, 4 debug

_DIVIDERS = {"large": "=" * 70, "medium": "=" * 42}


def print_divider(size):
    print(_DIVIDERS.get(size, "=" * 8))


def console(verbosity_level, module, msg):
    if verbosity_level <= VERBOSITY:
        print(module + " - " + msg)
