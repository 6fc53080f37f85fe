"""ArchiveStats record (API contract of ribs/archives/_archive_stats.py:10-47)."""
import dataclasses

import numpy as np


@dataclasses.dataclass
class ArchiveStats:
    """Aggregate statistics; floating members are 0-d arrays of the objective dtype."""

    num_elites: int
    coverage: np.floating
    qd_score: np.floating
    norm_qd_score: np.floating
    obj_max: np.floating | None
    obj_mean: np.floating | None
