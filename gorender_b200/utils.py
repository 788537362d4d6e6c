"""utils.go mirror."""


def isPowerOfTwo(n: int) -> bool:
    """utils.go:5-7."""
    return (n & (n - 1)) == 0
