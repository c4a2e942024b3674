""" per-frame image / region helpers of the reference's video/analysis package that sit on
the filter -> segment path, backed by the B200 kernels """
