"""train_generator_imperceptible.py of the reference: the alternated step of train_generator.py plus a total-variation term on the
poisoned images in the G-step loss,

    loss = loss_ce + L2_weight * loss_l2 + tv_weight * total_variation(inputs_bd).mean() + clean_model_weight * clean_model_loss
                                                                       (train_generator_imperceptible.py:227-237, --tv_weight 0.01)

with `gauss_smooth = T.GaussianBlur(3, (0.1, 1))` fixed at module level (:52) instead of --kernel_size / --sigma, and one more
scalar ("TV Loss", :304) on the tensorboard writer.  get_model / eval / the checkpoint dict are the base trainer's (the files differ
only in those lines and in debugging image dumps).  Same engine, same captured graphs; the TV term is one fused kernel
(csrc/blend.cu tv_loss_k: loss partials + its gradient added to the gradient arriving at inputs_bd).

`kornia.losses.total_variation` (kornia 0.6.6, requirements.txt:12) is not in the reference tree nor in this image: the formula is
restated from its published source (sum over C, H, W of the absolute vertical and horizontal differences, per image) -- parity of
this one term is unpinned w.r.t. kornia itself; everything around it is pinned by tests/golden/step_imperceptible_b32x2.npz.
"""
from __future__ import annotations

from . import train_generator as _base
from .train_generator import create_targets_bd, get_model, low_freq  # noqa: F401  (same definitions, :29-50,78-116)


def _variant(opt):
    opt.variant = "imperceptible"
    opt.kernel_size, opt.sigma = 3, (0.1, 1.0)   # module-level gauss_smooth (:52)
    if not hasattr(opt, "tv_weight"):
        opt.tv_weight = 0.01
    return opt


def train(netC, optimizerC, schedulerC, netG, optimizerG, schedulerG, netF, clean_model, train_dl, tf_writer, epoch, opt):
    """train_generator_imperceptible.py:119-316"""
    return _base.train(netC, optimizerC, schedulerC, netG, optimizerG, schedulerG, netF, clean_model, train_dl, tf_writer, epoch,
                       _variant(opt))


def eval(netC, optimizerC, schedulerC, netG, optimizerG, schedulerG, netF, clean_model, test_dl, best_clean_acc, best_bd_acc,
         best_F_acc, best_clean_model_acc, best_clean_model_bd_ba, best_clean_model_bd_asr, tf_writer, epoch, opt):
    """train_generator_imperceptible.py:319-460 (create_inputs_bd = the base trainer's trigger pipeline)"""
    return _base.eval(netC, optimizerC, schedulerC, netG, optimizerG, schedulerG, netF, clean_model, test_dl, best_clean_acc,
                      best_bd_acc, best_F_acc, best_clean_model_acc, best_clean_model_bd_ba, best_clean_model_bd_asr, tf_writer,
                      epoch, _variant(opt))


def main(argv=None):
    """train_generator_imperceptible.py:463-601: the base driver with this module's train / eval"""
    return _base.main(argv, train_fn=train, eval_fn=eval)


if __name__ == "__main__":
    main()
