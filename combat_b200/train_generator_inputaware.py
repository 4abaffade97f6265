"""train_generator_inputaware.py of the reference: the alternated step of train_generator.py with a SECOND loader and a
cross-trigger term in the G-step loss (train_generator_inputaware.py:180-185,234-264),

    inputs_bd2 = blur(clamp(inputs + low_freq(netG(inputs2)) * noise_rate))      -- the trigger of ANOTHER image on this image
    loss = CE(netC(T(inputs_bd)), bd_targets) + cross_weight * CE(netC(T(inputs_bd2)), targets)
           + L2_weight * MSE(inputs_bd, inputs) + clean_model_weight * CE(clean_model(T(inputs_bd)), targets)

(--cross_weight 0.2), `gauss_smooth = T.GaussianBlur(3, (0.1, 1))` fixed at module level (:53), netG's optimiser at
`lr_C * 0.1` on netC's schedule (:120-127), no gradient-image loss, one more counter ("Cross Acc") in train and eval, and
`best_cross_acc` / `mask` / `pattern` in the checkpoint dict (:490-497).  Same engine and kernels as the base step: the generator
runs ONCE over `[inputs ; inputs2]` (InstanceNorm is per sample), the cross leg is one more eval-mode netC forward + input
gradient, and both trigger gradients go back through one generator backward (engine.AlternatedStep, `variant = "inputaware"`).
Pinned by tests/golden/step_inputaware_b32x2.npz, recorded from the unmodified reference variant.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

from . import config
from . import train_generator as _base
from .engine import N_LOSSES, make_plan
from .train_generator import _adopt_momentum, _bind_momentum, _engine_for, create_targets_bd, low_freq  # noqa: F401


def _variant(opt):
    opt.variant = "inputaware"
    opt.kernel_size, opt.sigma = 3, (0.1, 1.0)   # module-level gauss_smooth (:53)
    if not hasattr(opt, "cross_weight"):
        opt.cross_weight = 0.2
    return opt


def get_model(opt):
    """train_generator_inputaware.py:83-138: the base construction; netG's SGD at lr_C * 0.1 on netC's milestones (:120-127)."""
    netC, optimizerC, schedulerC, netG, _, _, netF, clean_model = _base.get_model(opt)
    optimizerG = torch.optim.SGD(netG.parameters(), opt.lr_C * 0.1, momentum=0.9, weight_decay=5e-4, nesterov=True)
    schedulerG = torch.optim.lr_scheduler.MultiStepLR(optimizerG, opt.schedulerC_milestones, opt.schedulerC_lambda)
    return netC, optimizerC, schedulerC, netG, optimizerG, schedulerG, netF, clean_model


def _host_labels(targets):
    return targets.cpu().numpy() if torch.is_tensor(targets) else np.asarray(targets)


def _pin(t):
    return t if t.is_cuda else t.pin_memory()


def train(netC, optimizerC, schedulerC, netG, optimizerG, schedulerG, netF, clean_model, train_dl, train_dl2, mask, pattern,
          tf_writer, epoch, opt):
    """train_generator_inputaware.py:141-335 (one epoch; `mask` / `pattern` are unused there too)."""
    print(" Train:")
    opt = _variant(opt)
    netC.train()
    eng = _engine_for(netC, clean_model, netG, netF, opt)
    _adopt_momentum(optimizerC, netC)
    _adopt_momentum(optimizerG, netG)
    eng.set_lr(optimizerC.param_groups[0]["lr"], optimizerG.param_groups[0]["lr"])
    use_graph = not getattr(opt, "no_graph", False)
    log_every = max(1, int(getattr(opt, "log_every", 50)))
    dev = netC.net.device
    tot = torch.zeros(16, dtype=torch.int64, device=dev)
    lsum = torch.zeros(N_LOSSES, dtype=torch.float64, device=dev)
    total_sample, n_batches, acc = 0, len(train_dl), {}
    for batch_idx, (inputs, targets), (inputs2, _t2) in zip(range(n_batches), train_dl, train_dl2):   # :181
        y_host = _host_labels(targets)
        plan = make_plan(y_host, opt, eng.with_metrics)
        out = eng.step(_pin(inputs), y_host, plan, use_graph=use_graph, x2=_pin(inputs2))
        tot += out["counts"].long()
        lsum += out["losses"].double()
        total_sample += len(y_host)
        if (batch_idx + 1) % log_every == 0 or batch_idx + 1 == n_batches:
            c, l = tot.cpu().numpy(), lsum.cpu().numpy()
            acc = dict(avg_acc_clean=c[4] * 100.0 / total_sample, avg_acc_bd=c[6] * 100.0 / total_sample,
                       avg_acc_F=c[10] * 100.0 / total_sample, avg_acc_cross=c[12] * 100.0 / total_sample,
                       avg_clean_model_acc=c[2] * 100.0 / total_sample, avg_clean_model_bd_ba=c[8] * 100.0 / total_sample,
                       avg_clean_model_bd_asr=c[9] * 100.0 / total_sample, avg_loss_l2=l[2] / total_sample,
                       avg_clean_model_loss=l[3] / total_sample)
            print("[%d/%d] Clean Acc: %.4f | Bd Acc: %.4f | F Acc: %.4f | Cross Acc: %.4f | Clean Model Acc: %.4f | "
                  "Clean Model Bd BA: %.4f | Clean Model Bd ASR: %.4f"
                  % (batch_idx + 1, n_batches, acc["avg_acc_clean"], acc["avg_acc_bd"], acc["avg_acc_F"], acc["avg_acc_cross"],
                     acc["avg_clean_model_acc"], acc["avg_clean_model_bd_ba"], acc["avg_clean_model_bd_asr"]))
    if acc and not epoch % 1:                                                                  # :316-333
        tf_writer.add_scalars("Clean Accuracy", {
            "Clean": acc["avg_acc_clean"], "Bd": acc["avg_acc_bd"], "F": acc["avg_acc_F"], "Cross": acc["avg_acc_cross"],
            "CleanModel Acc": acc["avg_clean_model_acc"], "CleanModel Bd BA": acc["avg_clean_model_bd_ba"],
            "CleanModel Bd ASR": acc["avg_clean_model_bd_asr"], "L2 Loss": acc["avg_loss_l2"],
            "CleanModel Loss": acc["avg_clean_model_loss"]}, epoch)
    _bind_momentum(optimizerC, netC)
    _bind_momentum(optimizerG, netG)
    for n, b in netC.named_buffers():
        if n.endswith("num_batches_tracked"):
            b.fill_(netC.net.num_batches_tracked[n[: -len(".num_batches_tracked")]])
    schedulerC.step()
    schedulerG.step()


def eval(netC, optimizerC, schedulerC, netG, optimizerG, schedulerG, netF, clean_model, test_dl, test_dl2, mask, pattern,
         best_clean_acc, best_bd_acc, best_cross_acc, best_F_acc, best_clean_model_acc, best_clean_model_bd_ba,
         best_clean_model_bd_asr, tf_writer, epoch, opt):
    """train_generator_inputaware.py:338-506: the base evaluation plus the cross-trigger accuracy on the non-target rows
    (:402-413); saves when the clean accuracy improves (:471 -- no tie-break on the attack accuracy in this variant) with
    `best_cross_acc`, `mask`, `pattern` added to the checkpoint dict.  Returns the seven bests in the reference's order."""
    print(" Eval:")
    opt = _variant(opt)
    netC.eval()
    eng = _engine_for(netC, clean_model, netG, netF, opt)
    use_graph = not getattr(opt, "no_graph", False)
    dev = netC.net.device
    tot = torch.zeros(16, dtype=torch.int64, device=dev)
    n_clean = n_bd = 0
    for _, (inputs, targets), (inputs2, _t2) in zip(range(len(test_dl)), test_dl, test_dl2):
        out = eng.eval_step(_pin(inputs), _host_labels(targets), use_graph=use_graph, x2=_pin(inputs2))
        tot += out["counts"].long()
        n_clean += out["n"]
        n_bd += out["n_bd"]
    c = tot.cpu().numpy()
    n_bd_ = max(n_bd, 1)
    acc_clean, acc_bd, acc_cross, acc_F = (c[0] * 100.0 / n_clean, c[2] * 100.0 / n_bd_, c[12] * 100.0 / n_bd_, c[4] * 100.0 / n_bd_)
    acc_clean_model, bd_ba_clean_model, bd_asr_clean_model = c[6] * 100.0 / n_clean, c[8] * 100.0 / n_bd_, c[9] * 100.0 / n_bd_
    print("Clean Acc: {:.4f} - Best: {:.4f} | Bd Acc: {:.4f} - Best: {:.4f} | Cross Acc: {:.4f} - Best: {:.4f} | F Acc: {:.4f} - Best: "
          "{:.4f} | Clean Model BA: {:.4f} - Best: {:.4f} | Clean Model Bd BA: {:.4f} - Best: {:.4f} | Clean Model Bd ASR: {:.4f} - "
          "Best: {:.4f}".format(acc_clean, best_clean_acc, acc_bd, best_bd_acc, acc_cross, best_cross_acc, acc_F, best_F_acc,
                                acc_clean_model, best_clean_model_acc, bd_ba_clean_model, best_clean_model_bd_ba,
                                bd_asr_clean_model, best_clean_model_bd_asr))
    if not epoch % 1:
        tf_writer.add_scalars("Test Accuracy", {"Clean": acc_clean, "Bd": acc_bd, "Cross": acc_cross, "F": acc_F,
                                                "Clean Model Acc": acc_clean_model, "Clean Model Bd BA": bd_ba_clean_model,
                                                "Clean Model Bd ASR": bd_asr_clean_model}, epoch)
    if acc_clean > best_clean_acc:                                                             # :471
        print(" Saving...")
        best_clean_acc, best_bd_acc, best_cross_acc, best_F_acc = acc_clean, acc_bd, acc_cross, acc_F
        best_clean_model_acc, best_clean_model_bd_ba, best_clean_model_bd_asr = acc_clean_model, bd_ba_clean_model, bd_asr_clean_model
        state_dict = {
            "netC": netC.state_dict(), "schedulerC": schedulerC.state_dict(), "optimizerC": optimizerC.state_dict(),
            "netG": netG.state_dict(), "schedulerG": schedulerG.state_dict(), "optimizerG": optimizerG.state_dict(),
            "clean_model": clean_model.state_dict(), "best_clean_acc": acc_clean, "best_bd_acc": acc_bd,
            "best_cross_acc": acc_cross, "best_F_acc": acc_F, "best_clean_model_acc": best_clean_model_acc,
            "best_clean_model_bd_ba": best_clean_model_bd_ba, "best_clean_model_bd_asr": best_clean_model_bd_asr,
            "epoch_current": epoch, "mask": mask, "pattern": pattern,
        }
        ckpt_dir = os.path.dirname(opt.ckpt_path)
        if ckpt_dir:
            os.makedirs(ckpt_dir, exist_ok=True)
        torch.save(state_dict, opt.ckpt_path)
    return (best_clean_acc, best_bd_acc, best_cross_acc, best_F_acc, best_clean_model_acc, best_clean_model_bd_ba,
            best_clean_model_bd_asr)


def main(argv=None):
    """train_generator_inputaware.py:509-678: two train and two test loaders (:535-538, num_workers 0), get_model, the
    detector / clean-model checkpoints, --continue_training (+ best_cross_acc, mask, pattern), the (unused by the step)
    `mask` / `pattern` tensors (:615-617), n_iters epochs of train() + eval()."""
    import shutil
    opt = config.get_arguments().parse_args(argv)
    _base._dataset_shape(opt)
    opt.num_workers = 0                                                                       # :532
    from .utils.dataloader import get_dataloader
    train_dl, test_dl = get_dataloader(opt, True), get_dataloader(opt, False)
    train_dl2, test_dl2 = get_dataloader(opt, True), get_dataloader(opt, False)
    netC, optimizerC, schedulerC, netG, optimizerG, schedulerG, netF, clean_model = get_model(opt)
    mode = opt.saving_prefix
    opt.ckpt_folder = os.path.join(opt.checkpoints, "{}_clean".format(mode), opt.dataset)
    opt.ckpt_path = os.path.join(opt.ckpt_folder, "{}_{}_clean.pth.tar".format(opt.dataset, mode))
    opt.log_dir = os.path.join(opt.ckpt_folder, "log_dir")
    os.makedirs(opt.log_dir, exist_ok=True)
    if netF is not None:
        opt.F_ckpt_folder = os.path.join(opt.F_checkpoints, opt.dataset)
        opt.F_ckpt_path = os.path.join(opt.F_ckpt_folder, opt.F_model, "{}_{}_detector.pth.tar".format(opt.dataset, opt.F_model))
        if os.path.exists(opt.F_ckpt_path) or not opt.synthetic_data:
            print(f"Loading {opt.F_model} at {opt.F_ckpt_path}")
            netF.load_state_dict(torch.load(opt.F_ckpt_path, map_location=opt.device, weights_only=False)["netC"])
            print("Done")
        netF.eval()
    if opt.load_checkpoint_clean is not None or not opt.synthetic_data:
        load_path = os.path.join(opt.checkpoints, str(opt.load_checkpoint_clean), opt.dataset,
                                 "{}_{}.pth.tar".format(opt.dataset, opt.load_checkpoint_clean))
        if not os.path.exists(load_path):
            print("Error: {} not found".format(load_path))
            sys.exit()
        clean_model.load_state_dict(torch.load(load_path, map_location=opt.device, weights_only=False)["netC"])
    clean_model.eval()
    keys = ("best_clean_acc", "best_bd_acc", "best_cross_acc", "best_F_acc", "best_clean_model_acc", "best_clean_model_bd_ba",
            "best_clean_model_bd_asr")
    bests, epoch_current = [0.0] * 7, 0
    if opt.continue_training:
        if not os.path.exists(opt.ckpt_path):
            print("Pretrained model doesnt exist")
            sys.exit()
        print("Continue training!!")
        sd = torch.load(opt.ckpt_path, map_location=opt.device, weights_only=False)
        for name, obj in (("netC", netC), ("optimizerC", optimizerC), ("schedulerC", schedulerC), ("netG", netG),
                          ("optimizerG", optimizerG), ("schedulerG", schedulerG), ("clean_model", clean_model)):
            obj.load_state_dict(sd[name])
        bests = [sd[k] for k in keys]
        epoch_current = sd["epoch_current"]
        mask, pattern = sd["mask"], sd["pattern"]
    else:
        print("Train from scratch!!!")
        shutil.rmtree(opt.ckpt_folder, ignore_errors=True)
        os.makedirs(opt.log_dir, exist_ok=True)
        mask = torch.zeros(opt.input_height, opt.input_width).to(opt.device)                  # :615-617
        mask[2:6, 2:6] = 0.1
        pattern = torch.rand(opt.input_channel, opt.input_height, opt.input_width).to(opt.device)
    try:
        from torch.utils.tensorboard import SummaryWriter
        tf_writer = SummaryWriter(log_dir=opt.log_dir)
    except Exception:
        tf_writer = _base._NullWriter()
    for epoch in range(epoch_current, opt.n_iters):
        print("Epoch {}:".format(epoch + 1))
        train(netC, optimizerC, schedulerC, netG, optimizerG, schedulerG, netF, clean_model, train_dl, train_dl2, mask, pattern,
              tf_writer, epoch, opt)
        bests = list(eval(netC, optimizerC, schedulerC, netG, optimizerG, schedulerG, netF, clean_model, test_dl, test_dl2, mask,
                          pattern, *bests, tf_writer, epoch, opt))
    return bests


if __name__ == "__main__":
    main()
