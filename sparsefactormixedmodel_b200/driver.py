"""Host-side mirror of the reference's R driver contract (placeholder; filled in with the sweep)."""


class BSFGSweep:  # pragma: no cover - replaced in the next commit
    pass


def fast_BSFG_sampler(BSFG_state, n_samples):  # pragma: no cover
    raise NotImplementedError
