"""bench.py helpers that need no GPU: provenance of the ncu constants and the per-stage HBM view (CPU only)."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ncu_constants_describe_the_tree():
    import bench

    nc = bench.ncu_constants("celt")
    assert nc and nc["frames_per_step"] == 204800
    p = bench.ncu_provenance(nc)
    assert p["matches_this_source"], p  # the headline workload's capture is of the kernels in this tree
    ps = bench.ncu_provenance(bench.ncu_constants("silk"))
    assert ps["matches_this_source"] or ps.get("kernels_changed_since_capture"), ps  # ... or says which kernels changed


def test_stage_traffic_view():
    import bench

    nc = bench.ncu_constants("celt")
    line = json.loads(open(os.path.join(ROOT, "profiles", "r02_v6_bench_celt.json")).read().strip().splitlines()[-1])
    v = bench.stage_traffic_view(nc, 1.0, line["roofline"]["stage_ms"], 6554.2)
    assert set(v) >= {"entropy", "bands", "imdct", "postfilter", "deemph"}
    assert 0.6 < v["deemph"]["frac_of_hbm_peak"] < 1.0   # the de-emphasis is the HBM-bound stage
    assert all(x["frac_of_hbm_peak"] < 0.4 for k, x in v.items() if k != "deemph")
    assert bench.stage_traffic_view(None, 1.0, {"deemph": 1.0}, 6554.2) is None
    assert bench.stage_traffic_view(nc, None, {"deemph": 0.0}, 6554.2) is None
