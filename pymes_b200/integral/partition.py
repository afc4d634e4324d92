"""Named sub-blocks of the two-body integrals (reference pymes/integral/partition.py:4-39).

``part_2_body_int(no, V)`` returns the same 16 keys as the reference; every value
is a zero-copy strided *view* of V (numpy or torch), which the contraction engine
consumes directly through its stride descriptors."""

OCCUPIED = "ijkl"
KEYS = ("abci", "iabj", "iajk", "aijk", "klij", "aibj", "ijak", "abic", "iajb", "abcd", "iabc",
        "aijb", "ijka", "aibc", "ijab", "abij")


def block_slices(no, key):
    """Tuple of slices selecting block ``key`` (letters i-l occupied, a-d virtual)."""
    return tuple(slice(0, no) if ch in OCCUPIED else slice(no, None) for ch in key)


def part_2_body_int(no, t_V_pqrs):
    return {key: t_V_pqrs[block_slices(no, key)] for key in KEYS}
