"""Indented logging with the reference's semantics (pymes/log.py:4-32):
``level`` selects the indentation (4 spaces each) and messages with
``level > debug_level`` (default 3) are dropped.  Solver output lines keep the
reference's wording because users and tests grep them."""

INDENT = "    "
_state = {"debug_level": 3, "quiet": False}


def set_debug_level(n):
    """Global default verbosity (the reference hard-codes 3)."""
    _state["debug_level"] = int(n)


def set_quiet(flag=True):
    _state["quiet"] = bool(flag)


def print_logging_info(*args, level=0, debug_level=None, **_ignored):
    limit = _state["debug_level"] if debug_level is None else debug_level
    if _state["quiet"] or level > limit:
        return
    print(INDENT * level + "".join(map(str, args)))


def print_title(title_name, sep_symbol="=", level=1, debug_level=None):
    limit = _state["debug_level"] if debug_level is None else debug_level
    if _state["quiet"] or level > limit:
        return
    level = max(level, 1)
    width = 80 // level
    if width < len(title_name):
        width = len(title_name) + 2
    shift = (80 - width) // 2
    pad = (width - len(title_name)) // 2
    bar = " " * shift + sep_symbol * width
    print(bar)
    print(" " * (shift + pad) + title_name + " " * pad)
    print(bar)
