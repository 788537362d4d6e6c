"""utils.go mirror."""


def isPowerOfTwo(n: int) -> bool:
    """utils.go:7-9."""
    return n != 0 and (n & (n - 1)) == 0
