"""CPU-only: the oracle at the BASELINE size (300^3, 4096-ray batches, configs 2 and 3) against the
OUTPUT-ONLY goldens the unmodified reference produced (tests/golden/make_golden.py::gen_fullsize).
The 33 MB fields are regenerated from their seeds by the product's constructor, whose init equals
the reference's (test_host_logic.py); the parameter checksums stored with the goldens pin that."""
import numpy as np
import torch

import bench
from helpers import check_with_ties, load_golden
from oracle import oracle as orc

TOL = 1e-4


def scene300_cpu(n_fields):
    """bench.build_scene + the cfg-3 extension, as in make_golden.make_scene300."""
    lt = bench.build_scene("cpu", 300)
    for k in range(1, n_fields):
        lt.append_frame()
        torch.manual_seed(k)
        lt.append_rf(1)
    return lt


def oracle_fields(lt):
    out = []
    for rf in lt.tensorfs:
        fd = {k: v.detach().cpu().numpy() for k, v in rf.state_dict().items()}
        kw = rf.get_kwargs()
        for k in ("density_shift", "distance_scale", "rayMarch_weight_thres", "fea_pe", "view_pe",
                  "featureC", "app_dim", "step_ratio", "fea2denseAct", "gridSize"):
            fd[k] = kw[k]
        out.append(orc.Field(fd))
    return out


def checksums(lt):
    return np.array([float(p.detach().double().sum()) for rf in lt.tensorfs for p in rf.parameters()])


def test_cfg2_oracle_vs_reference_golden():
    g = load_golden("cfg2_300")
    lt = scene300_cpu(1)
    np.testing.assert_allclose(checksums(lt), g["param_checksum"], rtol=1e-12, atol=1e-12)  # sum order differs (channels_last)
    field = oracle_fields(lt)[0]
    for b in g["batches"]:
        ids = np.arange(b * 4096, (b + 1) * 4096, dtype=np.int64)
        out = bench.oracle_batch(lt, field, ids)
        n1, e1 = check_with_ties(out["rgb"], g[f"b{b}.rgb"], g[f"b{b}.margin"], TOL, f"cfg2 b{b} rgb")
        n2, e2 = check_with_ties(out["depth"], g[f"b{b}.depth"], g[f"b{b}.margin"], TOL, f"cfg2 b{b} depth")
        print(f"cfg2 batch {b}: rgb worst {e1:.2e} ({n1} threshold ties), depth worst {e2:.2e}")
        assert n2 == 0          # depth does not depend on the shading switch


def test_cfg3_oracle_vs_reference_golden():
    g = load_golden("cfg3_300")
    lt = scene300_cpu(3)
    np.testing.assert_allclose(checksums(lt), g["param_checksum"], rtol=1e-12, atol=1e-12)  # sum order differs (channels_last)
    fields = oracle_fields(lt)
    b = int(g["batch"])
    ids = np.arange(b * 4096, (b + 1) * 4096, dtype=np.int64)
    zs = [orc.sample_table(f.n_samples()) for f in fields]
    focal = float(lt.focal(800).detach())
    cx, cy = [float(v) for v in lt.center(800, 800).detach()]
    c2w = lt.get_cam2world(torch.tensor([0])).detach().numpy()
    expo = torch.stack(list(lt.exposure))[[0]].detach().numpy()
    out = orc.local_forward(fields, zs, ids, 800, 800, False, focal, cx, cy, c2w, g["world2rf"],
                            g["blend"], exposure=expo)
    n1, e1 = check_with_ties(out["rgb"], g["rgb"], g["margin"], TOL, "cfg3 rgb")
    n2, e2 = check_with_ties(out["depth"], g["depth"], g["margin"], TOL, "cfg3 depth")
    print(f"cfg3: rgb worst {e1:.2e} ({n1} threshold ties), depth worst {e2:.2e}")
    assert n2 == 0
