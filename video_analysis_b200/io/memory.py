"""
In-memory video (reference: video/io/memory.py:17-64): an ndarray (T, H, W[, 3]) seen
through the VideoBase protocol.  `get_frame` returns a view, as the reference does.

B200 additions (not in the reference): `pin()` page-locks the array so that batches
can be DMA'd to the device straight from it, and `frame_block(a, b)` hands the device
filters a contiguous (b-a, H, W[, 3]) slab without a per-frame Python loop.
"""

import numpy as np

from .base import VideoBase


class VideoMemory(VideoBase):
    write_access = True
    seekable = True

    def __init__(self, data, fps=25, copy_data=True):
        self.data = np.array(data, copy=True) if copy_data else np.asarray(data)
        # a trailing singleton colour axis is dropped (memory.py:33-34); unlike the
        # reference we then read the metadata from the squeezed array, so (T,H,W,1) works
        if self.data.ndim > 3 and self.data.shape[3] == 1:
            self.data = np.squeeze(self.data, 3)
        d = self.data
        if d.ndim == 3:
            is_color = False
        elif d.ndim == 4 and d.shape[3] == 3:
            is_color = True
        else:
            raise ValueError('The last dimension of the data must be either 1 or 3.')
        super(VideoMemory, self).__init__(size=(d.shape[2], d.shape[1]), frame_count=d.shape[0],
                                          fps=fps, is_color=is_color)
        self._pinned = False

    def get_frame(self, index):
        if index < 0:
            index += self.frame_count
        return self.data[index]

    def __getitem__(self, key):
        return self.data[key]

    def __setitem__(self, key, value):
        self.data[key] = value

    # ---- device-path helpers -------------------------------------------------------------
    def frame_block(self, start, stop):
        """ frames [start, stop) as one array (a view) """
        return self.data[start:stop]

    def pin(self):
        """ page-lock the backing array in place (cudaHostRegister) so that uploads are
        asynchronous DMA transfers; returns self """
        if not self._pinned and self.data.flags['C_CONTIGUOUS'] and self.data.nbytes:
            import torch
            rc = torch.cuda.cudart().cudaHostRegister(self.data.ctypes.data, self.data.nbytes, 0)
            if int(rc) != 0:
                raise RuntimeError('cudaHostRegister failed with code %d' % int(rc))
            self._pinned = True
        return self

    def close(self):
        if self._pinned:
            import torch
            torch.cuda.cudart().cudaHostUnregister(self.data.ctypes.data)
            self._pinned = False
