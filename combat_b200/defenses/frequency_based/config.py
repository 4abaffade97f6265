"""defenses/frequency_based/config.py of the reference (:4-32): the detector scripts' own parser -- same flags and defaults; the
build-only additions are the ones of combat_b200.config (--synthetic_data for the loader, no network here)."""
import argparse


def get_arguments():
    parser = argparse.ArgumentParser()
    parser.add_argument("--data_root", type=str, default="../../data/")
    parser.add_argument("--checkpoints", type=str, default="./checkpoints")
    parser.add_argument("--device", type=str, default="cuda")
    parser.add_argument("--results", type=str, default="./results")
    parser.add_argument("--dataset", type=str, default="cifar10")
    parser.add_argument("--model", type=str, default="original")
    parser.add_argument("--input_height", type=int, default=32)
    parser.add_argument("--input_width", type=int, default=32)
    parser.add_argument("--input_channel", type=int, default=3)
    parser.add_argument("--num_classes", type=int, default=10)
    parser.add_argument("--num_workers", type=int, default=2)
    parser.add_argument("--attack_mode", type=str, default="all2one")
    parser.add_argument("--n_iters", type=int, default=50)
    parser.add_argument("--bs", type=int, default=64)
    parser.add_argument("--noise_rate", type=float, default=0.08)
    parser.add_argument("--ratio", type=float, default=0.65, help="scale ratio for DCT of noise")
    parser.add_argument("--kernel_size", type=int, default=3, help="kernel size for Gaussian blur")
    parser.add_argument("--sigma", default=(0.1, 1.0), help="sigma for Gaussian blur")
    parser.add_argument("--continue_training", action="store_true")
    parser.add_argument("--load_checkpoint", default="../../checkpoints")
    parser.add_argument("--saving_prefix", type=str, help="Folder in /checkpoints for saving ckpt")
    parser.add_argument("--scale_noise_rate", type=float, default=1.0)
    parser.add_argument("--debug", action="store_true", default=False)
    parser.add_argument("--synthetic_data", action="store_true", default=False)   # build-only (see combat_b200.config)
    return parser
