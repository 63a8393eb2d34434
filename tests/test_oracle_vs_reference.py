"""Live cross-check of the oracle against the UNMODIFIED reference classes on seeds and action mixes other than those of
the committed golden traces.  Runs only where the reference tree exists (the build container); skipped on the GPU box,
where /root/reference is absent and nothing may read it."""
import numpy as np
import pytest

from oracle import dmp_oracle as O
from oracle import refload

pytestmark = pytest.mark.skipif(not refload.available(), reason="reference tree not present (GPU box)")

REF3D_P = [.2, .2, .2, .2, .05, .05, .05, .05]

CASES = [
    # dim, kind, ctor kwargs, seeds, steps, action probabilities (None = uniform)
    ("1D", "static", dict(plan_choose=1), (11, 12), 900, None),
    ("1D", "dynamic", dict(density="dense", split="val"), (13,), 900, [.2, .2, .6]),
    ("2D", "static", dict(plan_choose=1), (21,), 800, None),
    ("2D", "dynamic", dict(density="sparse", split="val"), (22, 23), 800, [.1, .2, .2, .1, .4]),
    ("3D", "static", dict(plan_choose=0), (31, 32), 800, REF3D_P),
    ("3D", "dynamic", dict(density="dense", split="test"), (33,), 800, None),
    ("3D", "dynamic", dict(density="sparse", split="train"), (34,), 800, [.1, .1, .1, .1, .15, .15, .15, .15]),
]


@pytest.mark.parametrize("dim,kind,kw,seeds,T,p", CASES)
def test_oracle_follows_live_reference(dim, kind, kw, seeds, T, p):
    d = int(dim[0])
    dynamic = kind == "dynamic"
    cls = refload.load_class(dim, kind)
    A = O.SPEC[d]["actions"]
    for seed in seeds:
        np.random.seed(seed)
        if dynamic:
            ref = cls(data_path=refload.dataset_path(dim, kw["density"], kw["split"]))
            orc = O.make_env(d, True, plans=[np.asarray(x, dtype=np.float64) for x in ref.plan_dataset])
        else:
            ref = cls(plan_choose=kw["plan_choose"])
            orc = O.make_env(d, False, plan_choose=kw["plan_choose"])
        arng = np.random.RandomState(seed + 500)

        def ref_obs(o):
            """(raw row, normalised row or None) of a reference observation in any of its formats."""
            if not dynamic:
                return np.asarray(o, dtype=np.float64).reshape(-1), None
            if d == 1:                                       # [raw, normalised, plan] (+ position on reset)
                return np.asarray(o[0], dtype=np.float64).reshape(-1), np.asarray(o[1], dtype=np.float64).reshape(-1)
            return None, np.asarray(o[0], dtype=np.float64).reshape(-1)       # [normalised, input_plan, position]

        def check_obs(o):
            rw, nrm = ref_obs(o)
            if rw is not None:
                assert np.array_equal(rw, orc.obs()[0])
            if nrm is not None:
                assert np.array_equal(nrm, orc.obs_normalised()[0])

        def reset_both():
            o = ref.reset()
            orc.reset(int(ref.index_random) if dynamic else 0)
            check_obs(o)
            assert float(ref.total_brick) == orc.total_brick
            assert np.array_equal(np.asarray(ref.plan, dtype=np.float64), orc.plan)

        reset_both()
        episodes = 0
        for t in range(T):
            a = int(arng.randint(A)) if p is None else int(arng.choice(A, p=p))
            o, r, dn = ref.step(a)
            _, rr, dd = orc.step(a, int(ref.step_size))      # the reference's own draw of this step
            check_obs(o)
            assert r == rr and isinstance(r, int) == isinstance(rr, int) and dn == dd, (seed, t, r, rr, dn, dd)
            assert np.array_equal(np.asarray(ref.environment_memory, dtype=np.float64), orc.grid), (seed, t)
            pos = ref.position_memory[-1]
            assert (pos == orc.pos) if d == 1 else (list(pos) == list(orc.pos))
            if d != 2:
                ri, oi = ref.iou(), orc.iou()
                assert ri == oi or (np.isnan(ri) and np.isnan(oi))
            if dn:
                episodes += 1
                reset_both()
        assert episodes >= 1
