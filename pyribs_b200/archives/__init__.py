"""Archives of the B200-native hot path (drop-in for `ribs.archives` on this path)."""
from pyribs_b200.archives._add_status import AddStatus
from pyribs_b200.archives._archive_base import DeviceArchive
from pyribs_b200.archives._archive_stats import ArchiveStats
from pyribs_b200.archives._categorical_archive import CategoricalArchive
from pyribs_b200.archives._cvt_archive import CVTArchive
from pyribs_b200.archives._device_store import DeviceStore
from pyribs_b200.archives._grid_archive import GridArchive
from pyribs_b200.archives._k_means import k_means_centroids

__all__ = ["AddStatus", "ArchiveStats", "CategoricalArchive", "CVTArchive", "DeviceArchive", "DeviceStore", "GridArchive", "k_means_centroids"]
