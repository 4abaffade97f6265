"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol the header declares, the CLI
parser matches the reference's flags/defaults, and the modules expose the reference's state_dict keys and shapes
(tests/golden/api.json was recorded from the unmodified reference by tests/golden/make_golden.py)."""
import json
import os
import re

import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_library_loads_and_exports_every_header_symbol():
    from combat_b200 import _lib
    declared = _lib.header_symbols()
    assert len(declared) >= 40
    for name in declared:
        assert hasattr(_lib.lib, name), "symbol %s declared in include/combat_b200.h is not exported" % name
        assert name in _lib._SIGS, "symbol %s has no ctypes signature" % name
    assert sorted(_lib._SIGS) == declared
    assert _lib.lib.combat_version() >= 100


def test_no_cpu_fallback():
    from combat_b200 import ops
    from combat_b200.nets import Classifier
    with pytest.raises(RuntimeError):
        ops.plane_op(torch.zeros(1, 3, 32, 32), "dct")
    with pytest.raises(RuntimeError):
        Classifier("preact_resnet18", device="cpu")


def test_product_code_never_imports_the_oracle():
    root = os.path.join(os.path.dirname(GOLDEN), "..", "combat_b200")
    for dp, _, fs in os.walk(root):
        for f in fs:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+\S*oracle", src, flags=re.M), os.path.join(dp, f)
                assert "combat_oracle" not in src and "import_module" not in src and "__import__" not in src, os.path.join(dp, f)


def test_cli_flags_match_reference():
    from combat_b200 import config
    g = json.load(open(os.path.join(GOLDEN, "api.json")))["flags"]
    p = config.get_arguments()
    mine = {a.dest: a for a in p._actions if a.dest != "help"}
    for dest, spec in g.items():
        assert dest in mine, dest
        d = mine[dest].default
        d = list(d) if isinstance(d, (list, tuple)) else d
        assert d == spec["default"], (dest, d, spec["default"])
    extra = set(mine) - set(g)
    assert extra == {"dtype", "no_graph", "log_every", "synthetic_data"}, extra
    opt = p.parse_args(["--pc", "0.3", "--noise_rate", "0.1", "--post_transform_option", "no_use"])
    assert opt.pc == 0.3 and opt.noise_rate == 0.1


@pytest.mark.parametrize("key,ctor", [
    ("PreActResNet18", lambda N: N.Classifier("preact_resnet18", 10, 3, 32, device="meta", dtype=torch.float32)),
    ("ResNet18_c8_64", lambda N: N.Classifier("resnet18", 8, 3, 64, device="meta", dtype=torch.float32)),
    ("UnetGenerator", lambda N: N.Generator(3, 64, 0, device="meta", dtype=torch.float32)),
    ("CUnetGeneratorv1_c8", lambda N: N.Generator(3, 64, 8, device="meta", dtype=torch.float32)),
    ("GridGenerator_s2", lambda N: N.GridGenerator(3, 64, 2, device="meta", dtype=torch.float32)),
])
def test_parameter_names_and_shapes_match_reference(key, ctor):
    from combat_b200 import nets
    g = json.load(open(os.path.join(GOLDEN, "api.json")))
    net = ctor(nets)
    ref = g["state_dicts"][key]
    ref_params = {k: v for k, v in ref.items() if not k.endswith(("running_mean", "running_var", "num_batches_tracked"))}
    assert list(net.store.names) == list(k for k in ref if k in ref_params) or set(net.store.names) == set(ref_params)
    for n in net.store.names:
        assert list(net.store.shapes[n]) == ref_params[n], n
    n_params = sum(int(torch.tensor(net.store.shapes[n]).prod()) for n in net.store.names)
    assert n_params == g["n_params"][key]
    if hasattr(net, "bns"):
        bufs = {b.name + s for b in net.bns for s in (".running_mean", ".running_var", ".num_batches_tracked")}
        assert bufs == set(ref) - set(ref_params)


@pytest.mark.parametrize("script", ["train_generator", "train_generator_multilabel", "train_generator_imperceptible",
                                    "train_generator_inputaware", "train_generator_wanet", "train_victim", "train_victim_multilabel",
                                    "train_victim_imperceptible", "train_victim_inputaware", "train_victim_wanet",
                                    "train_clean_classifier", "eval"])
def test_mirror_signatures_match_the_reference_scripts(script):
    """Every reference script with a mirror: get_model / train / eval / main (and the helpers a caller may import) exist under
    the same names and take the reference's positional parameters, in order (tests/golden/api.json holds the signatures parsed
    from the unmodified reference); mirrors may only APPEND optional parameters (e.g. main(argv=None))."""
    import importlib
    import inspect
    ref = json.load(open(os.path.join(GOLDEN, "api.json")))["signatures"][script]
    mod = importlib.import_module("combat_b200." + script)
    helpers = {"ViT", "vit_small", "create_inputs_bd", "create_inputs_bd_from_noise", "create_dir"}   # classifier zoo (out of
    #            scope) / inlined into the engine / os.makedirs
    for name, args in ref.items():
        if name in helpers:
            continue
        assert hasattr(mod, name), (script, name)
        prm = list(inspect.signature(getattr(mod, name)).parameters.values())
        mine = [p.name for p in prm]
        assert mine[:len(args)] == args, (script, name, mine, args)
        assert all(p.default is not inspect.Parameter.empty for p in prm[len(args):]), (script, name, mine)


def test_plan_matches_oracle_selection():
    """host-side poison selection / RNG order (engine.make_plan) == oracle.select_poison on the same seeds"""
    import numpy as np

    from combat_b200.engine import default_opt, make_plan
    from oracle import combat_oracle as O
    for seed in range(5):
        y = torch.randint(0, 10, (257,), generator=torch.Generator().manual_seed(seed))
        np.random.seed(seed)
        torch.manual_seed(seed)
        bd = O.create_targets_bd(y, "all2one", 0, 10)
        trg, ntrg, nbd = O.select_poison(y, bd, 0.5)
        s_c = O.draw_sigma() if nbd > 0 else None
        s_g = O.draw_sigma()
        np.random.seed(seed)
        torch.manual_seed(seed)
        plan = make_plan(y.numpy(), default_opt())
        assert plan.num_bd == nbd and np.array_equal(plan.trg_ind, trg.numpy()) and np.array_equal(plan.ntrg_ind, ntrg.numpy())
        assert plan.sigma_c == s_c and plan.sigma_g == s_g
        assert np.array_equal(plan.perm, torch.cat([trg, ntrg]).numpy())
    # empty and degenerate batches
    np.random.seed(0)
    plan = make_plan(np.array([3, 4, 5]), default_opt())  # no target-class row: nothing poisoned, no C-step sigma draw
    assert plan.num_bd == 0 and plan.sigma_c is None and list(plan.perm) == [0, 1, 2]
    plan = make_plan(np.array([1, 2, 3]), default_opt(attack_mode="all2all"))
    assert plan.num_bd == 0 and list(plan.bd_targets) == [2, 3, 4]
    with pytest.raises(Exception):
        make_plan(np.array([1]), default_opt(attack_mode="nope"))


def test_variant_plans_consume_the_rng_streams_like_the_reference_variants():
    """engine.make_plan for the input-aware step draws one more sigma right after the G-step's and a sixth transform between T3
    and T4 (train_generator_inputaware.py:234-242); for the WaNet step nothing but the poison count and the transforms
    (train_generator_wanet.py has no blur).  Checked against the oracle's own draws on the same seeds, transforms on."""
    import random

    import numpy as np

    from combat_b200.engine import default_opt, make_plan
    from oracle import combat_oracle as O
    y = torch.randint(0, 10, (40,), generator=torch.Generator().manual_seed(3))
    y[:5] = 0

    def seed():
        np.random.seed(9); torch.manual_seed(9); random.seed(9)

    for variant in ("inputaware", "wanet", ""):
        opt = default_opt(variant=variant, post_transform_option="use", dataset="cifar10")
        seed()
        bd = O.create_targets_bd(y, "all2one", 0, 10)
        _, _, nbd = O.select_poison(y, bd, 0.5)
        blur = variant != "wanet"
        s_c = O.draw_sigma() if (nbd > 0 and blur) else None
        t1, t2 = O.draw_post_transform(40, opt), O.draw_post_transform(40, opt)
        s_g = O.draw_sigma() if blur else None
        s_g2 = O.draw_sigma() if variant == "inputaware" else None
        t3 = O.draw_post_transform(40, opt)
        t6 = O.draw_post_transform(40, opt) if variant == "inputaware" else None
        t4, t5 = O.draw_post_transform(40, opt), O.draw_post_transform(40, opt)
        tail = float(torch.rand(1))
        seed()
        plan = make_plan(y.numpy(), opt)
        assert float(torch.rand(1)) == tail, variant          # the torch stream is at the same position afterwards
        assert plan.num_bd == nbd and plan.sigma_c == s_c and plan.sigma_g == s_g and plan.sigma_g2 == s_g2
        assert plan.tf.shape[0] == (6 if variant == "inputaware" else 5)
        for slot, prm in ((0, t1), (3, t2), (1, t3), (2, t4), (4, t5)) + (((5, t6),) if t6 is not None else ()):
            assert np.array_equal(plan.tf[slot][:, 0], (prm["xs"] - prm["pad"]).numpy().astype(np.float32)), (variant, slot)
            assert np.array_equal(plan.tf[slot][:, 5] != 0, prm["flip"].numpy()), (variant, slot)


def test_multilabel_plan_matches_oracle_chunks_and_draws():
    """engine.make_plan_multilabel == the oracle's restatement of train_generator_multilabel.py:171,203-221 on the same seeds:
    poison count from rand(bs), one sigma for the C-step (only if num_bd > 0), one per class chunk, chunk ids as targets."""
    import numpy as np

    from combat_b200 import ops
    from combat_b200.engine import default_opt, make_plan_multilabel, multilabel_chunks
    from oracle import combat_oracle as O
    for seed, (bs, ncls) in enumerate([(24, 10), (12, 8), (7, 10), (64, 8)]):
        y = torch.randint(0, ncls, (bs,), generator=torch.Generator().manual_seed(seed)).numpy()
        opt = default_opt(num_classes=ncls)
        np.random.seed(seed)
        torch.manual_seed(seed)
        nbd = int(np.sum(np.random.rand(bs) < opt.pc))
        s_c = O.draw_sigma() if nbd > 0 else None
        chunks = O.multilabel_chunks(bs, ncls)
        sig = [O.draw_sigma() for _ in chunks]
        np.random.seed(seed)
        torch.manual_seed(seed)
        plan = make_plan_multilabel(y, opt)
        assert multilabel_chunks(bs, ncls) == chunks
        assert plan.num_bd == nbd and plan.sigma_c == s_c and plan.sigmas_g == sig
        assert list(plan.perm) == list(range(bs)) and np.array_equal(plan.total_targets, y)   # rows keep their order and labels
        for (ci, si, ei), sg in zip(chunks, sig):
            assert (plan.bd_targets[si:ei] == ci).all()
            np.testing.assert_allclose(plan.taps_rows[si:ei], np.tile(ops.gaussian_taps(sg), (ei - si, 1)).astype(np.float32))
        assert sum(ei - si for _, si, ei in chunks) == bs


def test_param_store_padded_input_channels_round_trip():
    """ParamStore(store_ci=...): a conv weight stored with extra (zero) input channels keeps the reference's logical OIHW
    shape in p()/g()/state(), loads a reference tensor into the leading channels only, and lays the storage out
    channels-last with the stored width -- what the padded CUnetGeneratorv1.conv0_1 relies on."""
    import torch

    from combat_b200.nets import ParamStore
    specs = [("a.weight", (4, 5, 3, 3)), ("a.bias", (4,)), ("b.weight", (2, 3, 1, 1))]
    st = ParamStore(specs, "cpu", {"a.weight": 8})
    plain = ParamStore(specs, "cpu")
    assert st.shapes == plain.shapes and tuple(st.p("a.weight").shape) == (4, 5, 3, 3)
    assert st.raw(st.flat, "a.weight").numel() == 4 * 9 * 8 and plain.raw(plain.flat, "a.weight").numel() == 4 * 9 * 5
    assert st.offsets["a.bias"] == 4 * 9 * 8 and plain.offsets["a.bias"] == 4 * 9 * 5
    w = torch.randn(4, 5, 3, 3)
    sd = {"a.weight": w, "a.bias": torch.arange(4.0), "b.weight": torch.ones(2, 3, 1, 1)}
    st.load(sd)
    plain.load(sd)
    for k in sd:
        assert torch.equal(st.state()[k], sd[k]) and torch.equal(plain.state()[k], sd[k])
    stored = st.raw(st.flat, "a.weight").view(4, 3, 3, 8)
    assert torch.equal(stored[..., :5].permute(0, 3, 1, 2), w) and float(stored[..., 5:].abs().sum()) == 0.0
    st.g("a.weight").fill_(1.0)           # a gradient written through the logical view leaves the padding at zero
    assert float(st.raw(st.grad, "a.weight").sum()) == 4 * 5 * 9


def test_padded_conditional_generator_keeps_the_reference_shapes(monkeypatch):
    from combat_b200 import nets
    g = nets.Generator(num_classes=8, device="meta")
    monkeypatch.setenv("COMBAT_NO_PAD_COND", "1")
    ref = nets.Generator(num_classes=8, device="meta")
    monkeypatch.delenv("COMBAT_NO_PAD_COND")
    assert g.cond_pad and not ref.cond_pad
    assert g.convs["conv0_1"].Cin == 128 and ref.convs["conv0_1"].Cin == 72
    assert g.store.shapes == ref.store.shapes
    assert tuple(g.store.p("conv0_1.weight").shape) == (64, 72, 3, 3)
    assert g.store.numel - ref.store.numel == 64 * 9 * (128 - 72)


def test_instancenorm_split_policy_and_merge_formula():
    """Host side of the large-plane InstanceNorm (csrc/norm.cu): the split policy exported by the library (no GPU needed), and a
    numpy model of the merge the second kernel performs -- chunk (mean, centred sum of squares) partials combined with Chan's
    formula give the two-pass centred variance of the whole plane, ragged last chunk included."""
    import numpy as np

    from combat_b200._lib import lib
    f = lib.combat_instnorm_splits
    assert f(512, 1024, 64) == 1 and f(512, 256, 128) == 1          # CIFAR shapes: enough (sample, channel-group) pairs
    assert f(256, 4096, 64) == 3 and f(512, 16384, 64) == 1          # CelebA-size planes at batch 256: 256 CTAs -> 3 chunks
    k = f(64, 112 * 112, 64)
    assert 2 <= k <= 32 and (112 * 112) // k >= 256                  # ImageNet-10 shape, batch 64
    assert f(2, 72 * 72, 128) == 20                                  # capped by the 256-row minimum per chunk
    rng = np.random.default_rng(0)
    for HW, K in ((12544, 32), (5184, 20), (4096, 3)):
        x = rng.normal(40.0, 3.0, size=HW)
        rows = (HW + K - 1) // K
        parts = [(x[j * rows:(j + 1) * rows].mean(), ((x[j * rows:(j + 1) * rows] - x[j * rows:(j + 1) * rows].mean()) ** 2).sum(),
                  len(x[j * rows:(j + 1) * rows])) for j in range(K) if j * rows < HW]
        mean = sum(c * m for m, _, c in parts) / HW
        m2 = sum(q + c * (m - mean) ** 2 for m, q, c in parts)
        assert abs(mean - x.mean()) < 1e-12 and abs(m2 / HW - x.var()) < 1e-10 * x.var()


def test_utils_helpers_keep_the_reference_signatures(capsys):
    """utils/utils.py names the reference scripts import (`from utils.utils import progress_bar`)."""
    import inspect

    from combat_b200.utils import utils as U
    assert list(inspect.signature(U.progress_bar).parameters) == ["current", "total", "msg"]
    for i in range(3):
        U.progress_bar(i, 3, "Clean Acc: 1.0")
    out = capsys.readouterr().out
    assert out.endswith("3/3 \n") and out.count("\r") == 2 and "Clean Acc" in out
    assert U.format_time(0.0) == "0ms" and U.format_time(3.25) == "3s250ms" and U.format_time(3661) == "1h1m"
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3), torch.nn.BatchNorm2d(4), torch.nn.Flatten(), torch.nn.Linear(4, 2))
    U.init_params(net)
    assert float(net[1].weight.min()) == 1.0 and float(net[3].weight.abs().max()) < 1e-2


def test_inference_loader_yields_indices():
    """utils/dataloader_infer.py: (input, target, index) items; the synthetic batches carry the running sample index."""
    import types

    from combat_b200.utils import dataloader_infer as DI
    ds = DI.PoisonedDataset([(torch.zeros(3, 4, 4), 1), (torch.ones(3, 4, 4), 2)], 10, None)
    assert len(ds) == 2 and ds[1][1] == 2 and ds[1][2] == 1
    opt = types.SimpleNamespace(synthetic_data=True, debug=True, bs=8, input_channel=3, input_height=32, input_width=32, num_classes=10)
    dl = DI.get_dataloader(opt, train=False)
    x, y, idx = next(iter(dl))
    assert tuple(x.shape) == (8, 3, 32, 32) and idx.tolist() == list(range(8)) and len(dl) == 2


def test_parameter_block_layout_and_victim_plans_for_the_variants():
    """Host logic of the per-iteration parameter block (engine.AlternatedStep._plan_layout): every section 16-byte aligned, the
    PostTensorTransform section sized for five slots (six for the input-aware step), views that tile the block without overlap;
    make_plan_victim draws a blur sigma only for the additive trigger (the WaNet victim step draws nothing)."""
    import numpy as np

    from combat_b200.engine import TF_SLOT, AlternatedStep, default_opt, make_plan_victim
    from combat_b200.utils.dataloader import PARAM_WIDTH
    for B, ml, tf_on, n_tf in ((512, False, True, 5), (32, False, True, 6), (30, True, False, 5), (7, False, False, 5)):
        lay, nbytes = AlternatedStep._plan_layout(B, ml, tf_on, n_tf)
        spans = sorted((o, o + n) for o, n in lay.values() if n)
        assert all(o % 16 == 0 for o, _ in spans) and all(a[1] <= b[0] for a, b in zip(spans, spans[1:])) and spans[-1][1] <= nbytes
        assert lay["tf"][1] == (n_tf * B * PARAM_WIDTH * 4 if tf_on else 0) and lay["taps_rows"][1] == (8 * B if ml else 0)
        v = AlternatedStep._plan_views(torch.zeros(nbytes, dtype=torch.uint8), lay, B, n_tf)
        assert v["perm"].numel() == B and v["small"].numel() == 8 and (v["tf"] is None) == (not tf_on)
        if tf_on:
            assert tuple(v["tf"].shape) == (n_tf, B, PARAM_WIDTH)
    assert TF_SLOT["T6"] == 5 and sorted(TF_SLOT.values()) == list(range(6))
    y = np.array([0, 3, 0, 5, 0, 7])
    flags = np.array([True, False, True, False, False, False])
    for variant, draws in (("", True), ("wanet", False)):
        opt = default_opt(variant=variant)
        torch.manual_seed(4)
        tail = float(torch.rand(1)) if not draws else None
        torch.manual_seed(4)
        plan = make_plan_victim(y, flags, opt)
        assert plan.num_bd == 2 and list(plan.perm) == [0, 2, 1, 3, 4, 5] and (plan.sigma_c is not None) == draws
        if not draws:
            assert float(torch.rand(1)) == tail     # the torch stream was not touched
