"""Host-side placement for the PCIe-bound end-to-end pass: pin the calling process to the CPUs NVML reports as
local to its GPU, so that the pinned staging buffers it allocates afterwards (first-touch policy) and the copy
engine's reads stay on the GPU's NUMA node.  With one process per GPU on an 8-GPU box, eight H2D streams that all
pull from one socket's memory are what bounds `finetune()`; this is the usual `numactl --cpunodebind` done from
inside the process.  Best effort: any failure leaves the affinity untouched."""
from __future__ import annotations

import os
from typing import Dict


def bind_to_gpu(device_index: int) -> Dict:
    info = {"bound": False, "cpus": None, "reason": None}
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        if visible:                     # NVML enumerates physical devices: map the logical index through the mask
            ids = [v.strip() for v in visible.split(",") if v.strip()]
            tok = ids[device_index]
            handle = (pynvml.nvmlDeviceGetHandleByUUID(tok.encode() if isinstance(tok, str) else tok)
                      if tok.startswith("GPU-") or tok.startswith("MIG-") else pynvml.nvmlDeviceGetHandleByIndex(int(tok)))
        else:
            handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (ncpu + 63) // 64)
        local = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        target = local & allowed
        if not target:
            info["reason"] = "no NVML-local CPU is in this process's allowed set"
        elif target == allowed:
            info.update(bound=True, cpus=len(target), reason="already local")
        else:
            os.sched_setaffinity(0, target)
            info.update(bound=True, cpus=len(target), reason="restricted %d -> %d CPUs" % (len(allowed), len(target)),
                        previous=sorted(allowed))
    except Exception as exc:            # no NVML, no permission, ...: run unbound
        info["reason"] = "%s: %s" % (type(exc).__name__, exc)
    return info


def unbind(info: Dict) -> None:
    """Give the process back the CPU set it had before `bind_to_gpu` (for host-only work such as a CPU baseline)."""
    prev = info.pop("previous", None)
    if prev:
        try:
            os.sched_setaffinity(0, set(prev))
        except Exception:
            pass
