"""The INTEGRATION.md modules resolve through the reference's OWN registries inside a scratch copy of the unmodified
reference package (baseline/_ref/RL, or /root/reference/RL in the build container).  CPU only: class resolution and the
reference's error behaviour; construction needs a GPU (tests/test_gpu_trainer.py runs the reference trainer loop there)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_b200_modules_resolve_through_the_reference_registries():
    import reference_dropin as rd
    if rd.reference_package() is None:
        pytest.skip("no reference package on this host (baseline/_ref/RL is installed by __graft_entry__.build())")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "reference_dropin.py"), "registries"], capture_output=True,
                         text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    out = json.loads(res.stdout.strip().splitlines()[-1])
    for k in ("sampler", "buffer", "buffer_indexed", "algorithm", "approx_container", "trainer", "reference_ids_still_registered",
              "unknown_id_raises"):
        assert out[k] is True, (k, out)


def test_trainer_tags_match_reference_table():
    """The TensorBoard tag table is a naming contract with the reference's CSV / plotting tools
    (RL/utils/tensorboard_setup.py:13-40): identical keys and values."""
    import reference_dropin as rd
    src = rd.reference_package()
    if src is None:
        pytest.skip("no reference package on this host")
    code = ("import sys, json; sys.path.insert(0, %r); sys.path.insert(0, %r); import numpy as np\n"
            "np.float_ = getattr(np, 'float_', np.float64)\n"
            "from RL.utils.tensorboard_setup import tb_tags; print(json.dumps(tb_tags))" % (
                os.path.join(ROOT, "tests", "golden", "_gym_stub"), os.path.dirname(src)))
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stderr[-2000:]
    import msacl_b200  # noqa: F401
    from msacl_b200.trainer import tb_tags
    assert json.loads(res.stdout.strip().splitlines()[-1]) == tb_tags
