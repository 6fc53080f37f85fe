"""Test-infrastructure shim for python-box (used only by the reference's test fixtures)."""


class Box(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v
