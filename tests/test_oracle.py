"""CPU: pin the numpy oracle against every golden vector minted from the reference module."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLD, load_checkpoint
from oracle import tip_oracle as O

TOL = 5e-5   # fp32 noise floor module-vs-restatement measured in the survey: <= 7.7e-6


def _weights(g):
    if "checkpoint" in g.files:
        sd = load_checkpoint(str(g["checkpoint"]))
        if sd is None:
            pytest.skip("baseline/_ref checkpoint not staged")
        return sd, dict(with_rnn=True)
    return (O.random_state_dict(int(g["wseed"]), size_s=int(g["size_s"]),
                                with_rnn=bool(g["with_rnn"]), with_acc_sum=bool(g["with_acc_sum"])),
            dict(with_rnn=bool(g["with_rnn"])))


FIXTURES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLD, "*.npz"))
                  if "stream" not in p and "b256" not in p and "runner" not in p)


def test_fixture_inventory():
    assert len(FIXTURES) >= 14


@pytest.mark.parametrize("name", FIXTURES)
def test_oracle_matches_reference_golden(name):
    g = np.load(os.path.join(GOLD, name))
    sd, kw = _weights(g)
    keep = g["keep_mask"] if "keep_mask" in g.files else None
    scale = float(g["past_scale"]) if "past_scale" in g.files else 1.0
    y = O.forward(sd, g["x_imu"], g["x_s"], keep_mask=keep, past_scale=scale, **kw)
    assert y.shape == g["y"].shape and y.dtype == np.float32
    assert np.isfinite(y).all()
    err = np.abs(y - g["y"]).max()
    assert err < TOL, err


def test_oracle_fp64_budget():
    g = np.load(os.path.join(GOLD, "rw_b3_l39.npz"))
    sd, kw = _weights(g)
    y64 = O.forward(sd, g["x_imu"], g["x_s"], dtype=np.float64, **kw)
    assert np.abs(y64 - g["y"]).max() < TOL


def test_oracle_b256_subsample():
    g = np.load(os.path.join(GOLD, "rw_b256_l40_sub.npz"))
    sd = O.random_state_dict(int(g["wseed"]))
    x_imu, x_s = O.synth_inputs(int(g["xseed"]), 256, 40)
    y = O.forward(sd, x_imu, x_s)
    assert np.abs(y[g["idx"]] - g["y_sub"]).max() < TOL
    assert np.abs(y[:, -1] - g["y_last"]).max() < TOL
    assert abs(y.astype(np.float64).sum() - float(g["y_sum"])) < 1e-2


def test_oracle_b256_subsample_released_checkpoint():
    """BASELINE configs[1] with the released weights the bench uses (golden minted by oracle/make_golden_ck_b256.py)."""
    g = np.load(os.path.join(GOLD, "ck_b256_l40_sub.npz"))
    sd = load_checkpoint(str(g["checkpoint"]))
    if sd is None:
        pytest.skip("baseline/_ref checkpoint not staged")
    x_imu, x_s = O.synth_inputs(int(g["xseed"]), 256, 40)
    idx = g["idx"]
    y = O.forward(sd, x_imu[idx], x_s[idx])
    assert np.abs(y - g["y_sub"]).max() < TOL
    assert np.abs(y[:, -1] - g["y_last"][idx]).max() < TOL


def test_oracle_inputs_not_mutated_and_nan_handled():
    sd = O.random_state_dict(11)
    x_imu, x_s = O.synth_inputs(9, 2, 12, nan_frac=0.5)
    xi0, xs0 = x_imu.copy(), x_s.copy()
    y = O.forward(sd, x_imu, x_s)
    assert np.isnan(xs0).any() and np.isfinite(y).all()
    np.testing.assert_array_equal(x_imu, xi0)
    np.testing.assert_array_equal(np.isnan(x_s), np.isnan(xs0))


def test_oracle_stream_trace():
    g = np.load(os.path.join(GOLD, "ck_stream200.npz"))
    sd = load_checkpoint(str(g["checkpoint"]))
    if sd is None:
        pytest.skip("baseline/_ref checkpoint not staged")
    for t in (0, 1, 5, 38, 39, 40, 41, 120, 199):
        lo = max(0, t + 1 - 40)
        y = O.forward(sd, g["imu_rows"][lo:t + 1][None], g["s_rows"][lo:t + 1][None])
        assert np.abs(y[0, -1] - g["y_last"][t]).max() < TOL


def test_stochastic_generator_statistics():
    """The product's dropout-mask generator as the oracle restates it (include/tip_b200.h, tip_dropout):
    drop rates p = 0.8 (past state, reference :77) and 0.1 (nn.TransformerEncoderLayer default) within 4 sigma,
    kept elements scaled by 1/(1-p) (x5 resp. x1.11), distinct sites / seeds give independent masks, and the
    mean over masks is the identity (nn.Dropout is unbiased)."""
    n = 1 << 20
    idx = np.arange(n)
    for p, site in ((0.8, O.SEED_PAST), (0.1, O.seed_out(0)), (0.1, O.seed_ff1(3)), (0.1, O.seed_attn(2))):
        f = O.dropout_factors(1234567 + site, idx, p)
        drop = (f == 0).mean()
        assert abs(drop - p) < 4 * np.sqrt(p * (1 - p) / n) + 2e-5, (p, drop)
        kept = f[f != 0]
        assert np.all(kept == np.float32(1.0) / (np.float32(1.0) - np.float32(p)))
        assert abs(kept[0] - 1 / (1 - p)) < 1e-6 * 5
    a = O.dropout_factors(99 + O.SEED_PAST, idx, 0.8) == 0
    b = O.dropout_factors(100 + O.SEED_PAST, idx, 0.8) == 0
    c = O.dropout_factors(99 + O.seed_out(0), idx, 0.8) == 0
    for u, v in ((a, b), (a, c)):
        both = (u & v).mean()
        assert abs(both - 0.64) < 5e-3                                  # independent masks
    # consecutive elements (the four 16-bit lanes of one hash) are independent too
    assert abs((a[0::4] & a[1::4]).mean() - 0.64) < 5e-3 and abs((a[2::4] & a[3::4]).mean() - 0.64) < 5e-3
    x = np.random.RandomState(0).standard_normal(4096).astype(np.float32)
    acc = np.zeros(4096)
    for s in range(400):
        acc += x * O.dropout_factors(s * 7919 + O.SEED_PAST, np.arange(4096), 0.8)
    err = np.abs(acc / 400 - x)
    assert err.max() < 6 * np.abs(x).max() * 2.0 / np.sqrt(400)        # sd of one draw = 2|x| at p = 0.8


def test_stochastic_oracle_reduces_to_deterministic_and_is_seeded():
    sd = O.random_state_dict(5)
    x_imu, x_s = O.synth_inputs(6, 2, 40)
    y0 = O.forward(sd, x_imu, x_s)
    np.testing.assert_array_equal(O.forward(sd, x_imu, x_s, dropout=dict(seed=3)), y0)
    dp = dict(seed=3, past_state_dropout=0.8, encoder_dropout=0.1)
    y1, y2 = O.forward(sd, x_imu, x_s, dropout=dp), O.forward(sd, x_imu, x_s, dropout=dp)
    np.testing.assert_array_equal(y1, y2)
    y3 = O.forward(sd, x_imu, x_s, dropout=dict(dp, seed=4))
    assert np.abs(y1 - y0).max() > 1e-2 and np.abs(y1 - y3).max() > 1e-2 and np.isfinite(y1).all()
    # an explicit keep-mask equal to the generator's past-state mask reproduces the drawn-mask forward (:77)
    kin_pad = 256
    rows = np.arange(2 * 40, dtype=np.uint64).reshape(2, 40, 1)
    idx = rows * np.uint64(kin_pad) + np.uint64(90) + np.arange(131, dtype=np.uint64)
    keep = (O.dropout_factors(3 + O.SEED_PAST, idx, 0.8) != 0).astype(np.float32)
    ya = O.forward(sd, x_imu, x_s, keep_mask=keep, past_scale=float(np.float32(1) / (np.float32(1) - np.float32(0.8))))
    yb = O.forward(sd, x_imu, x_s, dropout=dict(seed=3, past_state_dropout=0.8))
    assert np.abs(ya - yb).max() < 1e-5


def test_window_assembler_shapes_and_warmup():
    rs = np.random.RandomState(0)
    wa = O.WindowAssembler()
    outs = []
    for t in range(60):
        R = np.linalg.qr(rs.standard_normal((6, 3, 3)))[0].reshape(-1)
        outs.append(wa.push(np.concatenate((R, rs.standard_normal(18)))))
    assert all(o is None for o in outs[:5])          # Appendix B: first 5 calls make no model call
    assert outs[5].shape == (1, 90)
    assert outs[44].shape == (40, 90) and outs[59].shape == (40, 90)
    # rows already written are immutable: window t+1 rows[:-1] == window t rows[1:] (IMU part)
    np.testing.assert_allclose(outs[59][:-1, :72], outs[58][1:, :72], atol=1e-12)


@pytest.mark.parametrize("name", ["rw_b3_l39.npz", "rw_b2_l33_nornn.npz", "rw_b5_l40_s119.npz",
                                  "ck_model-with-dip9and10_b1_l40.npz"])
def test_torch_port_matches_reference_golden(name):
    """The CPU PyTorch port timed by bench.py's reference arm is pinned by the same goldens."""
    from oracle import tip_oracle_torch as OT
    g = np.load(os.path.join(GOLD, name))
    sd, kw = _weights(g)
    y = OT.forward(OT.to_torch_state(sd), g["x_imu"], g["x_s"], **kw).numpy()
    assert np.abs(y - g["y"]).max() < TOL


def test_product_synthetic_generators_match_the_oracle_copies():
    """bench.py's measured arm draws its weights / inputs from tip_b200.synthetic (no oracle import on the product
    side); the oracle keeps its own copy -- they must stay identical."""
    from tip_b200 import synthetic as S
    for seed, B, L in ((1, 3, 40), (9, 2, 7)):
        a, b = O.synth_inputs(seed, B, L), S.synth_inputs(seed, B, L)
        assert all(np.array_equal(x, y, equal_nan=True) for x, y in zip(a, b))
    for kw in (dict(), dict(size_s=119), dict(with_rnn=False), dict(with_acc_sum=False)):
        sa, sb = O.random_state_dict(4, **kw), S.random_state_dict(4, **kw)
        assert list(sa) == list(sb) and all(np.array_equal(sa[k], sb[k]) for k in sa)
    assert S.state_dict_keys() == O.state_dict_keys()


def test_bench_reference_arm_runs_on_cpu_and_prints_one_json_line():
    """`bench.py --impl reference` (the driver's reference arm) needs no GPU: one JSON line on stdout with the
    contract's keys, timed on the UNMODIFIED reference module when the install is staged (else the torch port)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1", "--batch", "8"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "imu_frames_per_sec_seq40_6imu" and j["unit"] == "frames/s"
    assert j["higher_is_better"] is True and j["value"] > 0 and j["n_gpus"] == 1
    staged = os.path.exists(os.path.join(root, "baseline", "_ref", "reference", "simple_transformer_with_state.py"))
    assert j["cpu_baseline"]["kind"] == ("reference" if staged else "port") and j["cpu_baseline"]["cores"] >= 1
    assert "BASELINE configs[1]" in j["config"]["workload"]
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["d2h_bytes_per_step"] == 0


def test_bench_reference_arm_only_rank0_works_under_torchrun_env():
    """Launched under torchrun (N > 1) rank 0 alone runs the reference arm; the other ranks exit 0 silently."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "1", "--batch", "8"], capture_output=True, text=True,
                       timeout=600, env=env)
    assert r.returncode == 0 and r.stdout.strip() == "", (r.stdout, r.stderr[-1000:])


def test_bench_lane_count_follows_the_step_count():
    """bench.py deals its K timed steps to the lanes round-robin: the lane count is adjusted so that it divides K."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    assert b.pick_lanes(20, 5) == 5 and b.pick_lanes(100, 5) == 5
    assert b.pick_lanes(8, 5) == 4 and b.pick_lanes(12, 5) == 4 and b.pick_lanes(18, 5) == 6 and b.pick_lanes(9, 5) == 3
    assert b.pick_lanes(7, 5) == 5 and b.pick_lanes(7, 1) == 1 and b.pick_lanes(1, 5) == 5
