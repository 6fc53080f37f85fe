"""AddStatus values returned by `add` (API contract of ribs/archives/_add_status.py:6-44)."""
from enum import IntEnum


class AddStatus(IntEnum):
    """NOT_ADDED < IMPROVE_EXISTING < NEW, so that statuses sort by how good the outcome is."""

    NOT_ADDED = 0
    IMPROVE_EXISTING = 1
    NEW = 2
