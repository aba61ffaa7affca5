"""Host logic of the fast_BSFG_sampler mirror (driver.py) without a GPU: the Posterior bookkeeping of
`save_posterior_samples` (BSFG_functions.R:375-408) and the burn / thin schedule of fast_BSFG_sampler.R:272-279, driven
through a stand-in for BSFGSweep that produces a deterministic "chain" (state = function of the iteration number)."""
import numpy as np


class FakeSweep:
    """Implements the part of BSFGSweep that driver.fast_BSFG_sampler uses; k grows by one factor at iteration 7."""
    n, r, pl, b, r2 = 6, 6, 4, 1, 0

    def __init__(self):
        self.i = 0
        self.k_max = 5
        self._post = None

    def _state(self, i):
        k = 3 if i < 7 else 4
        f = lambda rows, cols, s: (np.arange(rows * cols, dtype=float).reshape(rows, cols, order="F") + 1) * s + i
        return dict(Lambda=f(self.pl, k, 0.5), Lambda_prec=f(self.pl, k, 0.1), Plam=f(self.pl, k, 0.2), F=f(self.n, k, 1.0),
                    F_a=f(self.r, k, 2.0), F_h2=np.arange(k) / 10.0 + i, B=f(self.b, self.pl, 3.0), E_a=f(self.r, self.pl, 4.0),
                    resid_Y_prec=np.arange(self.pl) + 10.0 * i, E_a_prec=np.arange(self.pl) + 20.0 * i, delta=np.arange(k) + 1.0 * i,
                    tauh=np.arange(k) + 2.0 * i, W=np.zeros((0, self.pl)), W_prec=np.zeros(self.pl), nrun=i)

    def set_state(self, cs):
        self.i = int(cs.get("nrun", 0))

    def sweep(self, n):
        self.i += n

    def get_state(self, with_E_a=True):
        return self._state(self.i)

    def posterior_begin(self, cap):
        self._post = dict(cap=cap, samples=[])

    def collect(self, n_iter, burn, thin):
        for _ in range(n_iter):
            self.i += 1
            if self.i > burn and (self.i - burn) % thin == 0:
                assert len(self._post["samples"]) < self._post["cap"]
                self._post["samples"].append(((self.i - burn) // thin, self._state(self.i)))

    def posterior(self):
        S, km = len(self._post["samples"]), self.k_max
        out = dict(Lambda=np.zeros((self.pl * km, S)), F=np.zeros((self.n * km, S)), F_a=np.zeros((self.r * km, S)),
                   delta=np.zeros((km, S)), F_h2=np.zeros((km, S)), resid_Y_prec=np.zeros((self.pl, S)),
                   E_a_prec=np.zeros((self.pl, S)), W_prec=np.zeros((self.pl, S)), k_max=km,
                   sp_num=np.array([s for s, _ in self._post["samples"]]), k=np.array([st["F"].shape[1] for _, st in self._post["samples"]]))
        first = last = None
        B = np.zeros((self.b, self.pl)); E = np.zeros((self.r, self.pl))
        for j, (sp, st) in enumerate(self._post["samples"]):
            k = st["F"].shape[1]
            for nm, rows in (("Lambda", self.pl), ("F", self.n), ("F_a", self.r)):
                pad = np.zeros((rows, km)); pad[:, :k] = st[nm]
                out[nm][:, j] = pad.ravel(order="F")
            out["delta"][:k, j] = st["delta"]; out["F_h2"][:k, j] = st["F_h2"]
            out["resid_Y_prec"][:, j] = st["resid_Y_prec"]; out["E_a_prec"][:, j] = st["E_a_prec"]
            B = (B * (sp - 1) + st["B"]) / sp; E = (E * (sp - 1) + st["E_a"]) / sp        # the device recursion, seeded with 0
        out["B"], out["E_a"], out["W"] = B, E, np.zeros((0, self.pl))
        return out

    def close(self):
        pass


def _reference_posterior(fake, burn, thin, upto):
    """save_posterior_samples applied literally, sample by sample (BSFG_functions.R:375-408)."""
    P = {}
    sp = 0
    for i in range(1, upto + 1):
        if i > burn and (i - burn) % thin == 0:
            sp = (i - burn) // thin
            st = fake._state(i)
            for nm in ("Lambda", "F", "F_a", "delta", "F_h2", "resid_Y_prec", "E_a_prec"):
                vec = np.asarray(st[nm]).ravel(order="F")
                M = P.get(nm, np.zeros((vec.size, 0)))
                if M.shape[1] < sp:
                    M = np.column_stack([M, np.zeros((M.shape[0], sp - M.shape[1]))])
                if vec.size > M.shape[0]:                                   # "expand factor sample matrix if necessary", :384-391
                    M = np.vstack([M, np.zeros((vec.size - M.shape[0], M.shape[1]))])
                M[:vec.size, sp - 1] = vec
                P[nm] = M
            for nm in ("B", "E_a"):
                P[nm] = (P.get(nm, 0.0) * (sp - 1) + st[nm]) / sp          # :403-404
    return P


def test_fast_BSFG_sampler_bookkeeping_matches_save_posterior_samples():
    from sparsefactormixedmodel_b200.driver import fast_BSFG_sampler
    fake = FakeSweep()
    burn, thin = 2, 3                                    # samples at i = 5, 8, 11, 14 (k grows 3 -> 4 at i = 7)
    state = dict(run_parameters=dict(burn=burn, thin=thin), current_state=fake._state(0), Posterior={})
    out = fast_BSFG_sampler(state, 9, sweep=fake)        # i = 1..9: two samples
    assert out["current_state"]["nrun"] == 9 and state["current_state"]["nrun"] == 0
    want = _reference_posterior(FakeSweep(), burn, thin, 9)
    for nm in ("Lambda", "F", "F_a", "delta", "F_h2", "resid_Y_prec", "E_a_prec"):
        assert np.array_equal(out["Posterior"][nm], want[nm]), nm
    out2 = fast_BSFG_sampler(out, 6, sweep=fake)         # continuation: i = 10..15, samples 3 and 4
    want = _reference_posterior(FakeSweep(), burn, thin, 15)
    for nm in ("Lambda", "F", "F_a", "delta", "F_h2", "resid_Y_prec", "E_a_prec"):
        assert np.array_equal(out2["Posterior"][nm], want[nm]), nm
    for nm in ("B", "E_a"):                              # running means survive the restart (README.md:33)
        assert np.allclose(out2["Posterior"][nm], want[nm], rtol=1e-14, atol=0), nm
    assert out2["Posterior"]["Lambda"].shape == (FakeSweep.pl * 4, 4)       # grown to k = 4, zero padded for the k = 3 sample
    assert not out2["Posterior"]["Lambda"][FakeSweep.pl * 3:, 0].any()


def test_no_sample_inside_the_call_runs_plain_sweeps():
    from sparsefactormixedmodel_b200.driver import fast_BSFG_sampler
    fake = FakeSweep()
    state = dict(run_parameters=dict(burn=100, thin=10), current_state=fake._state(0), Posterior={})
    out = fast_BSFG_sampler(state, 7, sweep=fake)
    assert out["current_state"]["nrun"] == 7 and not out["Posterior"]
