"""Host logic of the zero-copy result pool (DeviceArchive._out_buffers): a pooled page-locked buffer may
only be reused once the caller dropped every array derived from it."""
import numpy as np
import torch

from pyribs_b200.archives import _archive_base as ab


def test_pool_entry_is_busy_while_any_view_is_alive():
    st, va = torch.empty(16, dtype=torch.int32), torch.empty(16, dtype=torch.float64)
    entry = [st, va, st.numpy(), va.numpy()]
    assert ab._pool_entry_free(entry)
    out = {"status": entry[2][0:16], "value": entry[3][4:12]}
    assert not ab._pool_entry_free(entry)
    sub = out["value"][1:3]  # a view of a view still pins the buffer
    del out
    assert not ab._pool_entry_free(entry)
    kept = np.array(sub)  # a copy does not
    del sub
    assert ab._pool_entry_free(entry)
    assert kept.shape == (2,)
