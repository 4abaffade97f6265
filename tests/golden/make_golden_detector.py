"""Golden fixture of the frequency-detector trainer, recorded from the UNMODIFIED reference.

Run in the build container only (needs /root/reference), in its OWN interpreter -- the defense ships modules named
`config`, `dataloader`, `model` and `train` that shadow the repository-level ones make_golden.py imports:
    python tests/golden/make_golden_detector.py
Writes tests/golden/detector_b8x2.npz: two iterations of defenses/frequency_based/train.py:train() on synthetic batches
(CPU, 8 threads), observed from outside (forward hooks on the reference's FrequencyModel, a recording wrapper around
torch.nn.functional.cross_entropy, state_dict / optimiser state afterwards) plus one eval() batch.

Stand-ins (documented in oracle/detector_oracle.py): `albumentations` is absent, so GaussNoise / RandomShadow are the
oracle's stand-in classes -- every other line that runs is the reference's.  The initial weights are the oracle's seeded
initialisation loaded into the reference module, so the fixture does not have to carry them.
"""
import os
import random
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
REF = "/root/reference"
sys.dont_write_bytecode = True
sys.path.insert(0, ROOT)
from oracle import combat_oracle as O  # noqa: E402
from oracle import detector_oracle as D  # noqa: E402

sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, "defenses", "frequency_based"))
alb = types.ModuleType("albumentations")
alb.GaussNoise, alb.RandomShadow = D.StandInGaussNoise, D.StandInRandomShadow
sys.modules["albumentations"] = alb
import classifier_models  # noqa: E402
from classifier_models.densenet import DenseNet121  # noqa: E402
from classifier_models.mobilenetv2 import MobileNetV2  # noqa: E402
from classifier_models.resnet import ResNet18  # noqa: E402
from classifier_models.vgg import VGG  # noqa: E402

for n, o in dict(VGG=VGG, DenseNet121=DenseNet121, MobileNetV2=MobileNetV2, ResNet18=ResNet18).items():
    setattr(classifier_models, n, o)
import config  # noqa: E402  (defenses/frequency_based/config.py)
import train as T  # noqa: E402  (defenses/frequency_based/train.py)

torch.set_num_threads(8)
SEED, B, ITERS = 11, 8, 2


class NullWriter:
    def add_scalars(self, *a, **k):
        pass

    def add_scalar(self, *a, **k):
        pass


def main():
    opt = config.get_arguments().parse_args([])
    opt.device = "cpu"
    netC, optimizerC = T.get_model(opt)
    gen = torch.Generator().manual_seed(SEED)
    p0, b0 = O.init_frequency_model_state(gen)
    netC.load_state_dict({**p0, **b0}, strict=False)
    xs = [torch.rand(B, 3, 32, 32, generator=gen) for _ in range(ITERS + 1)]
    rec = {"x_final": [], "preds": [], "loss": [], "y_final": []}
    netC.register_forward_pre_hook(lambda m, a: rec["x_final"].append(a[0].detach().clone()))
    netC.register_forward_hook(lambda m, a, out: rec["preds"].append(out.detach().clone()))
    ce = F.cross_entropy

    def ce_rec(inp, target, *a, **k):
        out = ce(inp, target, *a, **k)
        rec["loss"].append(out.detach().clone())
        rec["y_final"].append(target.detach().clone())
        return out

    torch.nn.functional.cross_entropy = ce_rec
    np.random.seed(SEED)
    random.seed(SEED)
    torch.manual_seed(SEED + 1)
    T.train(netC, optimizerC, [(x, torch.zeros(B, dtype=torch.long)) for x in xs[:ITERS]], NullWriter(), 0, opt)
    torch.nn.functional.cross_entropy = ce
    out = {"seed": SEED, "x": torch.stack(xs)}
    for i in range(ITERS):
        out["x_final%d" % i], out["y_final%d" % i] = rec["x_final"][i], rec["y_final"][i]
        out["preds%d" % i], out["loss%d" % i] = rec["preds"][i], rec["loss"][i]
    sd = netC.state_dict()
    names = [k for k in sd if not k.endswith("num_batches_tracked")]
    out["final_names"] = np.array(names)
    out["final_sum"] = np.array([float(sd[k].double().sum()) for k in names])
    out["final_l2"] = np.array([float(sd[k].double().norm()) for k in names])
    for k in ("conv1.weight", "conv1.bias", "bn1.weight", "bn1.running_mean", "bn1.running_var", "bn6.running_var", "linear6.weight",
              "linear6.bias"):
        out["final." + k] = sd[k]
    st = optimizerC.state[netC.linear6.weight]
    out["adadelta.linear6.weight.square_avg"], out["adadelta.linear6.weight.acc_delta"] = st["square_avg"], st["acc_delta"]
    out["num_batches_tracked"] = sd["bn1.num_batches_tracked"]
    # one eval() batch (train.py:230-253): eval-mode network, no shuffle; best_acc high enough that nothing is saved
    n_before = len(rec["preds"])
    opt.ckpt_path = os.devnull
    T.eval(netC, optimizerC, [(xs[ITERS], torch.zeros(B, dtype=torch.long))], 101.0, NullWriter(), 0, opt)
    out["eval_x_final"], out["eval_preds"] = rec["x_final"][n_before], rec["preds"][n_before]
    path = os.path.join(HERE, "detector_b8x2.npz")
    np.savez_compressed(path, **{k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in out.items()})
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024), "losses", [float(l) for l in rec["loss"]])


if __name__ == "__main__":
    main()
