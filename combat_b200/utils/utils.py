"""utils/utils.py of the reference (:14-127), the names its scripts import: `progress_bar(current, total, msg=None)` (:55-94, a
stdout progress line per batch), `format_time(seconds)` (:97-127), `get_mean_and_std(dataset)` (:14-27), `init_params(net)`
(:30-45).  Host-side conveniences, no device work.  The mirrored trainers print one line every --log_every iterations instead
of one per batch (a per-batch line forces a device synchronisation per iteration); `progress_bar` is kept for callers that
import it."""
from __future__ import annotations

import sys
import time

import torch

_BAR = 40
_t_begin = time.time()
_t_last = _t_begin


def format_time(seconds: float) -> str:
    """Up to two units, largest first: 'D', 'h', 'm', 's', 'ms' (reference :97-127)."""
    ms = int(round(seconds * 1000))
    parts = []
    for unit, size in (("D", 86400000), ("h", 3600000), ("m", 60000), ("s", 1000), ("ms", 1)):
        n, ms = divmod(ms, size)
        if n:
            parts.append("%d%s" % (n, unit))
    return "".join(parts[:2]) or "0ms"


def progress_bar(current: int, total: int, msg: str | None = None) -> None:
    """One carriage-returned line: bar, step time, total time, message, position (reference :55-94)."""
    global _t_begin, _t_last
    now = time.time()
    if current == 0:
        _t_begin = now
    step, _t_last = now - _t_last, now
    done = int(_BAR * (current + 1) / max(total, 1))
    line = " [%s>%s] Step: %s | Tot: %s%s %d/%d " % ("=" * done, "." * (_BAR - done), format_time(step), format_time(now - _t_begin),
                                                    (" | " + msg) if msg else "", current + 1, total)
    sys.stdout.write(line + ("\n" if current >= total - 1 else "\r"))
    sys.stdout.flush()


def get_mean_and_std(dataset):
    """Per-channel mean and mean-of-per-image std over a dataset of (image [3, H, W], label) pairs (reference :14-27)."""
    mean, std = torch.zeros(3), torch.zeros(3)
    loader = torch.utils.data.DataLoader(dataset, batch_size=1, shuffle=True, num_workers=2)
    print("==> Computing mean and std..")
    for inputs, _ in loader:
        mean += inputs[0].flatten(1).mean(1)
        std += inputs[0].flatten(1).std(1)
    return mean / len(dataset), std / len(dataset)


def init_params(net):
    """Kaiming-normal convs, unit BatchNorm, N(0, 1e-3) linear layers (reference :30-45) for plain torch modules."""
    import torch.nn as nn
    for m in net.modules():
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight, mode="fan_out")
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.BatchNorm2d):
            nn.init.constant_(m.weight, 1)
            nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.Linear):
            nn.init.normal_(m.weight, std=1e-3)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
