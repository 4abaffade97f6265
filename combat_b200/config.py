"""The reference's CLI surface (config.py:4-86 of the reference): `get_arguments()` returns an argparse parser accepting the
same flags with the same types and defaults -- pinned flag by flag against the reference by tests/golden/api.json
(tests/test_api_cpu.py) -- plus four build-only flags that rename nothing: --dtype, --no_graph, --log_every, --synthetic_data.

The flags are declared as typed tables (name -> default) rather than one call per flag; flags this package does not act on
(the other attack variants' knobs) are still accepted so that existing command lines keep parsing."""
import argparse

_STR = {
    "device": "cuda", "dataset": "cifar10", "data_root": "./data", "checkpoints": "./checkpoints", "temps": "./temps",
    "model": "default", "model_clean": "default", "F_model": "original", "scale_mode": "bicubic",
    "F_checkpoints": "./defenses/frequency_based/checkpoints",
    "saving_prefix": None, "load_checkpoint_clean": None,            # no default in the reference either
}
_UNTYPED = {"attack_mode": "all2one", "load_checkpoint": ""}        # declared without type= in the reference
_INT = {
    "input_height": 32, "input_width": 32, "input_channel": 3, "num_classes": 10, "bs": 128, "n_iters": 200,
    "num_workers": 6, "target_label": 0, "kernel_size": 3, "random_rotation": 10, "random_crop": 5, "s": 2, "S2": 8,
    "lnoise": 8, "F_num_ensemble": 3,
}
_FLOAT = {
    "lr_C": 1e-2, "lr_G": 1e-2, "lr_clean": 1e-2, "schedulerC_lambda": 0.1, "schedulerG_lambda": 0.1,
    "scheduler_clean_lambda": 0.1, "noise_rate": 0.08, "pc": 0.5, "ratio": 0.65, "L2_weight": 0.02,
    "clean_model_weight": 0.8, "tv_weight": 0.01, "cross_weight": 0.2, "cross_rate": 1, "lambda_cov": 1,
    "grid_rescale": 0.15, "scale": 1, "nearest": 0, "F_dropout": 0.5, "scale_noise_rate": 1.0, "r": 1 / 4, "scale_factor": 0.5,
}
# `type=list` / `type=tuple` in the reference: usable only through their defaults -- kept as declared
_SEQ = {"schedulerC_milestones": (list, [100, 150]), "schedulerG_milestones": (list, [100, 150]),
        "scheduler_clean_milestones": (list, [100, 150]), "sigma": (tuple, (0.1, 1.0))}
_SWITCH = {"continue_training": None, "clamp": None, "noise_only": False, "debug": False}   # store_true (default as declared)
_HELP = {"saving_prefix": "Folder in /checkpoints for saving ckpt", "ratio": "scale ratio for DCT of noise",
         "kernel_size": "kernel size for Gaussian blur", "sigma": "sigma for Gaussian blur"}


def get_arguments():
    p = argparse.ArgumentParser()

    def add(name, **kw):
        if name in _HELP:
            kw["help"] = _HELP[name]
        p.add_argument("--" + name, **kw)

    for table, typ in ((_STR, str), (_INT, int), (_FLOAT, float)):
        for name, default in table.items():
            add(name, type=typ, **({} if default is None else {"default": default}))
    for name, default in _UNTYPED.items():
        add(name, default=default)
    for name, (typ, default) in _SEQ.items():
        add(name, type=typ, default=default)
    for name, default in _SWITCH.items():
        add(name, action="store_true", **({} if default is None else {"default": default}))
    add("post_transform_option", type=str, default="use", choices=["use", "no_use", "use_modified"])
    # build-only
    p.add_argument("--dtype", type=str, default="bf16", choices=["bf16", "fp32"], help="activation storage / conv operand type")
    p.add_argument("--no_graph", action="store_true", default=False, help="do not capture the step in a CUDA graph")
    p.add_argument("--synthetic_data", action="store_true", default=False,
                   help="uniform-noise images / random labels instead of the dataset (no network for torchvision's download)")
    p.add_argument("--log_every", type=int, default=50, help="read the device-side metric counters back every N iterations")
    return p
