"""Generate the committed golden fixtures by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
Writes tests/golden/*.npz.  The reference has no tests or golden vectors of its own
(SURVEY.md section 4), so these outputs of the reference itself are the parity pins for
oracle/combat_oracle.py (tests/test_oracle_golden.py) and, through the oracle and
directly, for the CUDA path (tests/test_*_gpu.py).

Everything recorded here is observed from outside the reference code: forward hooks on
the reference's own nn.Modules, recording wrappers around torch.nn.functional losses,
and state_dicts before/after `train()`.
"""
import os
import random
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle.ref_loader import NullWriter, load_reference  # noqa: E402

tg, config = load_reference()
from classifier_models.preact_resnet import PreActResNet18  # noqa: E402
from classifier_models.resnet import ResNet18  # noqa: E402
from defenses.frequency_based.model import FrequencyModel  # noqa: E402
from networks.models import CUnetGeneratorv1, GridGenerator, UnetGenerator  # noqa: E402
from utils.dct import dct_2d, idct_2d  # noqa: E402

torch.set_num_threads(8)


def seed_all(s):
    torch.manual_seed(s)
    np.random.seed(s)
    random.seed(s)


def get_opt(extra=()):
    opt = config.get_arguments().parse_args(["--device", "cpu", "--post_transform_option", "no_use", *extra])
    return opt


def save(name, **kw):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **{k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in kw.items()})
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))


# ---------------------------------------------------------------- DCT
def gen_dct():
    seed_all(1)
    out = {}
    for N in (32, 64):
        x = torch.rand(2, 3, N, N) * 255
        xu = (torch.rand(2, 3, N, N) * 256).clamp(0, 255).byte()
        xn = torch.rand(2, 3, N, N) * 2 - 1
        opt = get_opt()
        opt.input_height = N
        out.update({
            "x%d" % N: x, "dct%d" % N: dct_2d(x), "idct%d" % N: idct_2d(x),
            "xu%d" % N: xu, "dctu%d" % N: dct_2d(xu),
            "xn%d" % N: xn, "lowfreq%d" % N: tg.low_freq(xn, opt),
        })
    # odd size (full-plane, non power of two) as used at ImageNet shape, kept tiny: N=28 plays the role of 224
    x = torch.rand(1, 3, 28, 28) * 255
    out.update(x28=x, dct28=dct_2d(x), idct28=idct_2d(x))
    save("dct.npz", **out)


# ---------------------------------------------------------------- modules
def tensor_digest(t):
    t = t.detach().flatten().double()
    return np.array([t.sum().item(), t.abs().sum().item(), (t * t).sum().item()])


def gen_modules():
    out = {}
    # UnetGenerator forward + input-grad + weight-grad digests, seed 2
    seed_all(2)
    opt = get_opt()
    netG = UnetGenerator(opt)
    x = (torch.rand(4, 3, 32, 32) * 2 - 1).requires_grad_(True)
    y = netG(x)
    w = torch.rand(4, 3, 32, 32)
    (y * w).sum().backward()
    out.update(unet_x=x, unet_y=y, unet_w=w, unet_dx=x.grad)
    for k, p in netG.named_parameters():
        out["unet_g_" + k] = tensor_digest(p.grad)
        out["unet_p_" + k] = tensor_digest(p)
    out["unet_gfull_conv0_0.weight"] = netG.conv0_0.weight.grad
    out["unet_gfull_upconv0_0.weight"] = netG.upconv0_0.weight.grad
    out["unet_gfull_conv3_1.bias"] = netG.conv3_1.bias.grad
    # CUnetGeneratorv1 (CelebA: 8 classes) at 64x64, seed 3
    seed_all(3)
    opt.num_classes = 8
    netG2 = CUnetGeneratorv1(opt)
    x2 = torch.rand(2, 3, 64, 64) * 2 - 1
    y2 = torch.randint(0, 8, (2,))
    out.update(cunet_x=x2, cunet_lbl=y2, cunet_y=netG2(x2, y2))
    # PreActResNet18: train-mode forward/backward + running stats, then eval-mode forward, seed 4
    seed_all(4)
    netC = PreActResNet18()
    x3 = (torch.rand(8, 3, 32, 32) * 2 - 1).requires_grad_(True)
    t3 = torch.randint(0, 10, (8,))
    netC.train()
    logits = netC(x3)
    loss = F.cross_entropy(logits, t3)
    loss.backward()
    out.update(preact_x=x3, preact_t=t3, preact_logits_train=logits, preact_loss=loss, preact_dx=x3.grad)
    for k, p in netC.named_parameters():
        out["preact_g_" + k] = tensor_digest(p.grad)
    out["preact_gfull_conv1.weight"] = netC.conv1.weight.grad
    out["preact_gfull_linear.weight"] = netC.linear.weight.grad
    out["preact_gfull_layer2.0.shortcut.0.weight"] = netC.layer2[0].shortcut[0].weight.grad
    out["preact_rm_layer1.0.bn1"] = netC.layer1[0].bn1.running_mean
    out["preact_rv_layer4.1.bn2"] = netC.layer4[1].bn2.running_var
    netC.eval()
    out["preact_logits_eval"] = netC(x3.detach())
    # ResNet18 (CelebA 64x64, 8 classes), seed 5
    seed_all(5)
    netR = ResNet18(num_classes=8)
    x4 = torch.rand(2, 3, 64, 64) * 2 - 1
    netR.train()
    out.update(resnet_x=x4, resnet_logits_train=netR(x4))
    netR.eval()
    out["resnet_logits_eval"] = netR(x4)
    # FrequencyModel with the SHIPPED cifar10 detector weights: dct_2d(uint8) -> netF known answer
    seed_all(6)
    netF = FrequencyModel(2, 3, 32)
    ck = torch.load(os.path.join("/root/reference/defenses/frequency_based/checkpoints/cifar10/cifar10_original_detector.pth.tar"),
                    map_location="cpu")
    sd = ck["netC"] if "netC" in ck else ck
    netF.load_state_dict(sd)
    netF.eval()
    xu = (torch.rand(4, 3, 32, 32) * 256).clamp(0, 255).byte()
    with torch.no_grad():
        out.update(freq_xu=xu, freq_logits=netF(dct_2d(xu)))
    # the shipped weights are 1.2 MB: keep them as fp16-free exact float32 in the fixture so the GPU box has them
    for k, v in netF.state_dict().items():
        out["freq_sd_" + k] = v
    save("modules.npz", **out)


# ---------------------------------------------------------------- full alternated step(s)
class Recorder:
    """Observes an unmodified train() call from outside."""

    def __init__(self, netC, netG, clean_model, netF):
        self.calls = {"netC": [], "netG": [], "clean": [], "netF": []}
        self.losses = []
        self.sigmas = []
        for name, m in (("netC", netC), ("netG", netG), ("clean", clean_model), ("netF", netF)):
            m.register_forward_hook(lambda mod, inp, outp, name=name: self.calls[name].append(
                (inp[0].detach().clone(), outp.detach().clone(), mod.training)))
        self._ce, self._mse = F.cross_entropy, F.mse_loss

    def __enter__(self):
        rec = self

        def ce(*a, **k):
            v = rec._ce(*a, **k)
            rec.losses.append(("ce", float(v)))
            return v

        def mse(*a, **k):
            v = rec._mse(*a, **k)
            rec.losses.append(("mse", float(v)))
            return v

        F.cross_entropy, F.mse_loss = ce, mse
        import torchvision.transforms as T
        self._gp = T.GaussianBlur.get_params

        def gp(lo, hi):
            s = rec._gp(lo, hi)
            rec.sigmas.append(s)
            return s

        T.GaussianBlur.get_params = staticmethod(gp)
        return self

    def __exit__(self, *a):
        F.cross_entropy, F.mse_loss = self._ce, self._mse
        import torchvision.transforms as T
        T.GaussianBlur.get_params = staticmethod(self._gp)


def param_summary(prefix, before, after, out, full_keys=()):
    for k in after:
        if not torch.is_floating_point(after[k]):
            out[prefix + "int_" + k] = after[k]
            continue
        d = (after[k] - before[k]).double()
        out[prefix + "dnorm_" + k] = np.array([d.norm().item(), d.sum().item()])
        out[prefix + "dhead_" + k] = (after[k] - before[k]).flatten()[:8]
        if k in full_keys:
            out[prefix + "dfull_" + k] = after[k] - before[k]


def gen_step(name, B, n_batches, seed, mod=None):
    """SURVEY 8c-4 recipe: seed all three RNGs, get_model (netC, clean_model, netG, netF), then data.
    mod: the reference module whose UNMODIFIED get_model / train run (default train_generator; the imperceptible variant passes
    train_generator_imperceptible, whose kornia.losses.total_variation is the stand-in of oracle/ref_loader.py)."""
    import tempfile
    tg = mod or globals()["tg"]
    opt = get_opt()
    opt.input_height = opt.input_width = 32
    opt.input_channel = 3
    os.chdir(tempfile.mkdtemp())   # the variant dumps a debugging image every 5 batches into the RELATIVE opt.temps (:280-285)
    seed_all(seed)
    netC, optC, schC, netG, optG, schG, netF, clean = tg.get_model(opt)
    netF.eval()
    clean.eval()
    batches = [(torch.rand(B, 3, 32, 32) * 2 - 1, torch.randint(0, 10, (B,))) for _ in range(n_batches)]
    sd0 = {n: {k: v.clone() for k, v in m.state_dict().items()} for n, m in (("netC", netC), ("netG", netG), ("clean", clean))}
    rec = Recorder(netC, netG, clean, netF)
    with rec:
        tg.train(netC, optC, schC, netG, optG, schG, netF, clean, batches, NullWriter(), 1, opt)
    torch.autograd.set_detect_anomaly(False)
    out = {"seed": seed, "B": B, "n_batches": n_batches, "sigmas": np.array(rec.sigmas)}
    out["loss_kinds"] = np.array([k for k, _ in rec.losses])
    out["loss_values"] = np.array([v for _, v in rec.losses])
    for i, (x, y) in enumerate(batches):
        out["y_%d" % i] = y
    # per iteration module calls: netG [C-step subset, G-step full]; netC [train total_x, eval x, eval x_bd];
    # clean [x, x_bd]; netF [dct]
    gi = ci = 0
    for i, (x, y) in enumerate(batches):
        g_sel_in, g_sel_out, _ = rec.calls["netG"][gi]
        g_all_in, g_all_out, g_train = rec.calls["netG"][gi + 1]
        gi += 2
        # recover which rows were poisoned by matching the generator's C-step input rows against the batch
        idx = [int((x == r).flatten(1).all(1).nonzero()[0, 0]) for r in g_sel_in]
        out["poison_idx_%d" % i] = np.array(idx, dtype=np.int64)
        out["num_bd_%d" % i] = len(idx)
        c_tot_in, c_tot_out, c_tr = rec.calls["netC"][ci]
        c_cl_in, c_cl_out, _ = rec.calls["netC"][ci + 1]
        c_bd_in, c_bd_out, _ = rec.calls["netC"][ci + 2]
        ci += 3
        assert c_tr and not g_train is False
        # batch order of total_x: recover the permutation
        perm = [int((x == r).flatten(1).all(1).nonzero()[0, 0]) if (x == r).flatten(1).all(1).any() else -1 for r in c_tot_in]
        out["total_perm_%d" % i] = np.array(perm, dtype=np.int64)  # -1 = poisoned (modified) rows
        out["logits_c_%d" % i] = c_tot_out
        out["pred_clean_%d" % i] = c_cl_out
        out["pred_bd_%d" % i] = c_bd_out
        out["clean_preds_%d" % i] = rec.calls["clean"][2 * i][1]
        out["clean_model_preds_%d" % i] = rec.calls["clean"][2 * i + 1][1]
        out["pred_F_%d" % i] = rec.calls["netF"][i][1]
        out["inputs_F_head_%d" % i] = rec.calls["netF"][i][0][:2]
        out["x_bd_head_%d" % i] = c_bd_in[:4]
        out["x_bd_digest_%d" % i] = tensor_digest(c_bd_in)
        out["noise_head_%d" % i] = g_all_out[:4]
        out["x_bd_c_%d" % i] = c_tot_in[: len(idx)]
    param_summary("netC_", sd0["netC"], netC.state_dict(), out, full_keys=("conv1.weight", "linear.weight", "linear.bias", "layer1.0.bn1.running_mean", "layer1.0.bn1.running_var", "layer3.1.bn2.weight"))
    param_summary("netG_", sd0["netG"], netG.state_dict(), out, full_keys=("conv0_0.weight", "conv0_0.bias", "upconv0_0.weight", "upconv0_0.bias", "conv3_1.bias"))
    param_summary("clean_", sd0["clean"], clean.state_dict(), out)
    save(name, **out)


def gen_step_inputaware(name, B, n_batches, seed):
    """Iterations of the UNMODIFIED train_generator_inputaware.train() (:141-335): two loaders (the second one supplies the images
    whose triggers are pasted on the first one's images), cross-trigger loss, lr_G = 0.1 * lr_C (get_model :120-127).
    Per iteration the reference calls netG 3x (C-step subset, x, x2), netC 4x (train total_x; eval x, inputs_bd2, inputs_bd),
    clean_model 2x, netF 1x; losses in call order: ce(C), ce(bd), ce(cross), mse, ce(clean model)."""
    import tempfile
    from oracle.ref_loader import load_reference_inputaware
    ti = load_reference_inputaware()
    opt = get_opt()
    opt.input_height = opt.input_width = 32
    opt.input_channel = 3
    os.chdir(tempfile.mkdtemp())   # debugging image dumps every 5 batches go to the RELATIVE opt.temps (:305-312)
    seed_all(seed)
    netC, optC, schC, netG, optG, schG, netF, clean = ti.get_model(opt)
    netF.eval()
    clean.eval()
    batches = [(torch.rand(B, 3, 32, 32) * 2 - 1, torch.randint(0, 10, (B,))) for _ in range(n_batches)]
    batches2 = [(torch.rand(B, 3, 32, 32) * 2 - 1, torch.randint(0, 10, (B,))) for _ in range(n_batches)]
    sd0 = {n: {k: v.clone() for k, v in m.state_dict().items()} for n, m in (("netC", netC), ("netG", netG), ("clean", clean))}
    rec = Recorder(netC, netG, clean, netF)
    with rec:
        ti.train(netC, optC, schC, netG, optG, schG, netF, clean, batches, batches2, None, None, NullWriter(), 1, opt)
    torch.autograd.set_detect_anomaly(False)
    out = {"seed": seed, "B": B, "n_batches": n_batches, "sigmas": np.array(rec.sigmas), "lr_G": optG.param_groups[0]["lr"]}
    out["loss_kinds"] = np.array([k for k, _ in rec.losses])
    out["loss_values"] = np.array([v for _, v in rec.losses])
    assert len(rec.calls["netG"]) == 3 * n_batches and len(rec.calls["netC"]) == 4 * n_batches
    for i, (x, y) in enumerate(batches):
        out["y_%d" % i] = y
        g_sel_in, _, _ = rec.calls["netG"][3 * i]
        _, g_all_out, _ = rec.calls["netG"][3 * i + 1]
        g2_in, g2_out, _ = rec.calls["netG"][3 * i + 2]
        assert torch.equal(g2_in, batches2[i][0])
        idx = [int((x == r).flatten(1).all(1).nonzero()[0, 0]) for r in g_sel_in]
        out["poison_idx_%d" % i] = np.array(idx, dtype=np.int64)
        out["num_bd_%d" % i] = len(idx)
        c_tot_in, c_tot_out, c_tr = rec.calls["netC"][4 * i]
        _, c_cl_out, _ = rec.calls["netC"][4 * i + 1]
        c_x_in, c_x_out, _ = rec.calls["netC"][4 * i + 2]
        c_bd_in, c_bd_out, _ = rec.calls["netC"][4 * i + 3]
        assert c_tr
        perm = [int((x == r).flatten(1).all(1).nonzero()[0, 0]) if (x == r).flatten(1).all(1).any() else -1 for r in c_tot_in]
        out["total_perm_%d" % i] = np.array(perm, dtype=np.int64)
        out["logits_c_%d" % i] = c_tot_out
        out["pred_clean_%d" % i] = c_cl_out
        out["pred_cross_%d" % i] = c_x_out
        out["pred_bd_%d" % i] = c_bd_out
        out["clean_preds_%d" % i] = rec.calls["clean"][2 * i][1]
        out["clean_model_preds_%d" % i] = rec.calls["clean"][2 * i + 1][1]
        out["pred_F_%d" % i] = rec.calls["netF"][i][1]
        out["x_bd_head_%d" % i] = c_bd_in[:4]
        out["x_bd2_head_%d" % i] = c_x_in[:4]
        out["x_bd2_digest_%d" % i] = tensor_digest(c_x_in)
        out["noise_head_%d" % i] = g_all_out[:4]
        out["noise2_head_%d" % i] = g2_out[:4]
        out["x_bd_c_%d" % i] = c_tot_in[: len(idx)]
    param_summary("netC_", sd0["netC"], netC.state_dict(), out, full_keys=("conv1.weight", "linear.weight", "linear.bias", "layer1.0.bn1.running_mean", "layer1.0.bn1.running_var", "layer3.1.bn2.weight"))
    param_summary("netG_", sd0["netG"], netG.state_dict(), out, full_keys=("conv0_0.weight", "conv0_0.bias", "upconv0_0.weight", "upconv0_0.bias", "conv3_1.bias"))
    param_summary("clean_", sd0["clean"], clean.state_dict(), out)
    save(name, **out)


def gen_step_wanet(name, B, n_batches, seed):
    """Iterations of the UNMODIFIED train_generator_wanet.train() (:92-300): GridGenerator, bicubic flow, grid_sample warp.
    Per iteration: netG 2x (C-step subset, x), netC 3x, clean_model 2x, netF 1x; losses in call order ce(C), ce(bd), mse (noise
    grid), mse, mse (logged gradient pair), ce(clean model).  No blur: no sigma draws.  The reference raises when an iteration
    poisons no row (reshape of an empty tensor, models.py:383): the seed is chosen so that every iteration has num_bd > 0."""
    import tempfile
    from oracle.ref_loader import load_reference_wanet
    tw = load_reference_wanet()
    opt = get_opt()
    opt.input_height = opt.input_width = 32
    opt.input_channel = 3
    os.chdir(tempfile.mkdtemp())
    seed_all(seed)
    netC, optC, schC, netG, optG, schG, netF, clean = tw.get_model(opt)
    netF.eval()
    clean.eval()
    batches = [(torch.rand(B, 3, 32, 32) * 2 - 1, torch.randint(0, 10, (B,))) for _ in range(n_batches)]
    sd0 = {n: {k: v.clone() for k, v in m.state_dict().items()} for n, m in (("netC", netC), ("netG", netG), ("clean", clean))}
    a = torch.linspace(-1, 1, steps=opt.input_height)   # :560-562
    gx, gy = torch.meshgrid(a, a)
    ident = torch.stack((gy, gx), 2)[None, ...]
    rec = Recorder(netC, netG, clean, netF)
    with rec:
        tw.train(netC, optC, schC, netG, optG, schG, netF, clean, batches, ident, NullWriter(), 1, opt)
    torch.autograd.set_detect_anomaly(False)
    out = {"seed": seed, "B": B, "n_batches": n_batches, "lr_G": optG.param_groups[0]["lr"], "s": opt.s,
           "grid_rescale": opt.grid_rescale}
    out["loss_kinds"] = np.array([k for k, _ in rec.losses])
    out["loss_values"] = np.array([v for _, v in rec.losses])
    assert len(rec.calls["netG"]) == 2 * n_batches and len(rec.calls["netC"]) == 3 * n_batches
    for i, (x, y) in enumerate(batches):
        out["y_%d" % i] = y
        g_sel_in, _, _ = rec.calls["netG"][2 * i]
        _, g_all_out, _ = rec.calls["netG"][2 * i + 1]
        idx = [int((x == r).flatten(1).all(1).nonzero()[0, 0]) for r in g_sel_in]
        out["poison_idx_%d" % i] = np.array(idx, dtype=np.int64)
        out["num_bd_%d" % i] = len(idx)
        c_tot_in, c_tot_out, c_tr = rec.calls["netC"][3 * i]
        _, c_cl_out, _ = rec.calls["netC"][3 * i + 1]
        c_bd_in, c_bd_out, _ = rec.calls["netC"][3 * i + 2]
        assert c_tr
        perm = [int((x == r).flatten(1).all(1).nonzero()[0, 0]) if (x == r).flatten(1).all(1).any() else -1 for r in c_tot_in]
        out["total_perm_%d" % i] = np.array(perm, dtype=np.int64)
        out["logits_c_%d" % i] = c_tot_out
        out["pred_clean_%d" % i] = c_cl_out
        out["pred_bd_%d" % i] = c_bd_out
        out["clean_preds_%d" % i] = rec.calls["clean"][2 * i][1]
        out["clean_model_preds_%d" % i] = rec.calls["clean"][2 * i + 1][1]
        out["pred_F_%d" % i] = rec.calls["netF"][i][1]
        out["x_bd_head_%d" % i] = c_bd_in[:8]
        out["x_bd_digest_%d" % i] = tensor_digest(c_bd_in)
        out["flow_%d" % i] = g_all_out
        out["x_bd_c_%d" % i] = c_tot_in[: len(idx)]
    param_summary("netC_", sd0["netC"], netC.state_dict(), out, full_keys=("conv1.weight", "linear.weight", "linear.bias", "layer1.0.bn1.running_mean", "layer1.0.bn1.running_var"))
    param_summary("netG_", sd0["netG"], netG.state_dict(), out, full_keys=("conv0_0.weight", "conv0_0.bias", "conv3_1.bias", "fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias"))
    save(name, **out)


def gen_mstep(name, dataset, B, n_batches, seed):
    """One or more iterations of the UNMODIFIED train_generator_multilabel.train() (reference :142-318).
    cifar10: models from the reference's own get_model; celeba: get_model passes an unknown keyword to CUnetGeneratorv1
    (TypeError as shipped), so the same classes are constructed here in get_model's order."""
    from oracle.ref_loader import load_reference_multilabel
    tm = load_reference_multilabel()
    opt = get_opt()
    opt.dataset = dataset
    if dataset == "cifar10":
        opt.input_height = opt.input_width = 32
        opt.input_channel, opt.num_classes = 3, 10
    else:
        opt.input_height = opt.input_width = 64
        opt.input_channel, opt.num_classes = 3, 8
    seed_all(seed)
    if dataset == "cifar10":
        netC, optC, schC, netG, optG, schG, netF, clean = tm.get_model(opt)
    else:
        netC, clean = ResNet18(num_classes=8), ResNet18(num_classes=8)
        netG = CUnetGeneratorv1(opt)
        netF = FrequencyModel(num_classes=2, n_input=3, input_size=64)
        optC = torch.optim.SGD(netC.parameters(), opt.lr_C, momentum=0.9, weight_decay=5e-4, nesterov=True)
        schC = torch.optim.lr_scheduler.MultiStepLR(optC, opt.schedulerC_milestones, opt.schedulerC_lambda)
        optG = torch.optim.SGD(netG.parameters(), opt.lr_C * 0.1, momentum=0.9, weight_decay=5e-4, nesterov=True)
        schG = torch.optim.lr_scheduler.MultiStepLR(optG, opt.schedulerC_milestones, opt.schedulerC_lambda)
    netF.eval()
    clean.eval()
    H = opt.input_height
    batches = [(torch.rand(B, 3, H, H) * 2 - 1, torch.randint(0, opt.num_classes, (B,))) for _ in range(n_batches)]
    sd0 = {n: {k: v.clone() for k, v in m.state_dict().items()} for n, m in (("netC", netC), ("netG", netG), ("clean", clean), ("netF", netF))}
    rec = Recorder(netC, netG, clean, netF)
    with rec:
        tm.train(netC, optC, schC, netG, optG, schG, netF, clean, batches, None, None, NullWriter(), 1, opt)
    torch.autograd.set_detect_anomaly(False)
    out = {"seed": seed, "B": B, "n_batches": n_batches, "sigmas": np.array(rec.sigmas), "num_classes": opt.num_classes}
    out["loss_kinds"] = np.array([k for k, _ in rec.losses])
    out["loss_values"] = np.array([v for _, v in rec.losses])
    # module calls per iteration: netG [C-step rows (skipped by the hook? no: called even when empty)] + one per class chunk;
    # netC [train total_x, eval x, eval x_bd]; clean [x, x_bd]; netF [dct]
    gi = 0
    for i, (x, y) in enumerate(batches):
        out["y_%d" % i] = y
        g_c_in = rec.calls["netG"][gi][0]
        out["num_bd_%d" % i] = g_c_in.shape[0]
        n_chunks = len([1 for ci in range(opt.num_classes) if ci * (int((B - 1) / opt.num_classes) + 1) < B])
        gi += 1 + n_chunks
        out["logits_c_%d" % i] = rec.calls["netC"][3 * i][1]
        out["total_x_head_%d" % i] = rec.calls["netC"][3 * i][0][:4]
        out["pred_clean_%d" % i] = rec.calls["netC"][3 * i + 1][1]
        out["pred_bd_%d" % i] = rec.calls["netC"][3 * i + 2][1]
        out["x_bd_%d" % i] = rec.calls["netC"][3 * i + 2][0]
        out["clean_preds_%d" % i] = rec.calls["clean"][2 * i][1]
        out["clean_model_preds_%d" % i] = rec.calls["clean"][2 * i + 1][1]
        out["pred_F_%d" % i] = rec.calls["netF"][i][1]
    # the initial weights are NOT stored (3 x 11 M floats): the oracle's init_* functions reproduce them from the same seed
    # and construction order; these digests pin that
    for n in ("netC", "netG", "clean", "netF"):
        out["init_digest_" + n] = tensor_digest(torch.cat([v.flatten().double() for k, v in sd0[n].items() if torch.is_floating_point(v)]))
    param_summary("netC_", sd0["netC"], netC.state_dict(), out, full_keys=("conv1.weight", "linear.weight", "linear.bias"))
    param_summary("netG_", sd0["netG"], netG.state_dict(), out, full_keys=("conv0_0.weight", "conv0_1.weight", "upconv0_0.bias"))
    save(name, **out)


def gen_eval(name, B, n_batches, seed):
    """The UNMODIFIED eval() of the reference (train_generator.py:321-465) over synthetic batches."""
    import tempfile
    opt = get_opt()
    opt.input_height = opt.input_width = 32
    opt.input_channel = 3
    opt.ckpt_path = os.path.join(tempfile.mkdtemp(), "ckpt.pth.tar")
    seed_all(seed)
    netC, optC, schC, netG, optG, schG, netF, clean = tg.get_model(opt)
    netF.eval()
    clean.eval()
    netG.eval()
    batches = [(torch.rand(B, 3, 32, 32) * 2 - 1, torch.randint(0, 10, (B,))) for _ in range(n_batches)]
    rec = Recorder(netC, netG, clean, netF)
    with rec:
        best = tg.eval(netC, optC, schC, netG, optG, schG, netF, clean, batches, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, NullWriter(), 1, opt)
    ck = torch.load(opt.ckpt_path, weights_only=False)
    out = {"seed": seed, "B": B, "n_batches": n_batches, "sigmas": np.array(rec.sigmas), "best": np.array([float(b) for b in best]),
           "ckpt_keys": np.array(sorted(ck.keys())), "ckpt_netC_keys": np.array(list(ck["netC"].keys()))}
    for i, (x, y) in enumerate(batches):
        out["y_%d" % i] = y
        out["preds_clean_%d" % i] = rec.calls["netC"][2 * i][1]
        out["preds_bd_%d" % i] = rec.calls["netC"][2 * i + 1][1]
        out["x_bd_%d" % i] = rec.calls["netC"][2 * i + 1][0]
        out["preds_F_%d" % i] = rec.calls["netF"][i][1]
        out["cm_clean_%d" % i] = rec.calls["clean"][2 * i][1]
        out["cm_bd_%d" % i] = rec.calls["clean"][2 * i + 1][1]
    save(name, **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["dct", "modules", "step"]
    if "eval" in which:
        gen_eval("eval_b64x2.npz", 64, 2, 5)
    if "mstep" in which:
        gen_mstep("mstep_cifar_b24x2.npz", "cifar10", 24, 2, 3)   # 10 classes, chunks of 3 (last chunk short: 24 = 7*3+3)
        gen_mstep("mstep_celeba_b12.npz", "celeba", 12, 1, 4)     # CelebA shape 64x64, 8 classes, ResNet18 + CUnetGeneratorv1
    if "dct" in which:
        gen_dct()
    if "modules" in which:
        gen_modules()
    if "step" in which:
        gen_step("step_b128.npz", 128, 1, 0)      # the SURVEY 8c-4 known-answer vector
        gen_step("step_b32x2.npz", 32, 2, 7)      # two iterations: momentum buffers + RNG interleaving
    if "variants" in which:
        from oracle.ref_loader import load_reference_imperceptible
        gen_step("step_imperceptible_b32x2.npz", 32, 2, 11, mod=load_reference_imperceptible())   # + tv_weight * TV(x_bd).mean()
    if "wanet" in which:
        gen_step_wanet("step_wanet_b32x2.npz", 32, 2, int(os.environ.get("WANET_SEED", "17")))
    if "inputaware" in which:
        gen_step_inputaware("step_inputaware_b32x2.npz", 32, 2, 13)   # second loader + cross-trigger loss


def gen_api():
    """API surface pins: parser flags/defaults and state_dict key -> shape maps of the reference modules."""
    import json
    p = config.get_arguments()
    flags = {}
    for a in p._actions:
        if a.dest == "help":
            continue
        d = a.default
        flags[a.dest] = {"default": list(d) if isinstance(d, (list, tuple)) else d,
                         "type": getattr(a.type, "__name__", None) if a.type else ("bool" if a.nargs == 0 else None)}
    opt = get_opt()
    opt8 = get_opt()
    opt8.num_classes = 8
    mods = {
        "PreActResNet18": PreActResNet18(), "ResNet18_c8_64": ResNet18(num_classes=8),
        "UnetGenerator": UnetGenerator(opt), "CUnetGeneratorv1_c8": CUnetGeneratorv1(opt8), "FrequencyModel": FrequencyModel(2, 3, 32),
        "GridGenerator_s2": GridGenerator(opt),
    }
    sds = {k: {n: list(v.shape) for n, v in m.state_dict().items()} for k, m in mods.items()}
    nparams = {k: sum(p.numel() for p in m.parameters()) for k, m in mods.items()}
    # positional signatures of the top-level functions of every reference script that has a mirror (parsed, not imported)
    import ast
    sigs = {}
    for script in ("train_generator", "train_generator_multilabel", "train_generator_imperceptible", "train_generator_inputaware",
                   "train_generator_wanet", "train_victim", "train_victim_multilabel", "train_victim_imperceptible",
                   "train_victim_inputaware", "train_victim_wanet", "train_clean_classifier", "eval"):
        tree = ast.parse(open(os.path.join("/root/reference", script + ".py")).read())
        sigs[script] = {n.name: [a.arg for a in n.args.args] for n in tree.body if isinstance(n, ast.FunctionDef)}
    with open(os.path.join(HERE, "api.json"), "w") as f:
        json.dump({"flags": flags, "state_dicts": sds, "n_params": nparams, "signatures": sigs}, f, indent=0, sort_keys=True)
    print("wrote api.json")


if __name__ == "__main__" and "api" in (sys.argv[1:] or ["api"]):
    gen_api()
