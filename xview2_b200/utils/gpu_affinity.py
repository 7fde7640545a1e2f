"""``set_affinity(gpu_id, mode)`` with the reference's modes (/root/reference/utils/gpu_affinity.py:126-146): pins the rank's
process (hence its loader decode threads) to CPU cores local to its GPU, as NVML reports them.

Restated compactly: NVML's per-GPU CPU mask -> (drop hyper-thread siblings) -> split the cores of a socket among the GPUs
attached to it ("socket_unique_*") -> add the siblings back.  Any NVML / sysfs failure leaves the affinity untouched
(the reference would raise; a loader that cannot be pinned still has to run)."""
import collections
import glob
import math
import os


def _gpu_cpu_list(nvml, index):
    words = math.ceil((os.cpu_count() or 1) / 64)
    handle = nvml.nvmlDeviceGetHandleByIndex(index)
    cores = []
    for w, mask in enumerate(nvml.nvmlDeviceGetCpuAffinity(handle, words)):
        cores += [64 * w + b for b in range(64) if (mask >> b) & 1]
    return cores


def _thread_siblings():
    """{first hardware thread: its sibling} from sysfs."""
    pairs = {}
    for path in glob.glob("/sys/devices/system/cpu/cpu*/topology/thread_siblings_list"):
        try:
            with open(path) as f:
                ids = [int(t) for t in f.read().strip().replace("-", ",").split(",") if t]
        except (OSError, ValueError):
            continue
        if len(ids) >= 2:
            pairs[ids[0]] = ids[1]
    return pairs


def _socket_unique(nvml, gpu_id, world_size, interleaved):
    siblings = _thread_siblings()
    drop = set(siblings.values())
    by_socket = collections.defaultdict(list)
    for dev in range(world_size):
        cores = tuple(sorted(set(_gpu_cpu_list(nvml, dev)) - drop))
        by_socket[cores].append(dev)
    for cores, devs in by_socket.items():
        if gpu_id not in devs or not cores:
            continue
        slot, n = devs.index(gpu_id), len(devs)
        per = max(1, len(cores) // n)
        mine = list(cores[slot::n]) if interleaved else list(cores[slot * per:(slot + 1) * per])
        return mine + [siblings[c] for c in mine if c in siblings]
    return None


def set_affinity(gpu_id=None, mode="socket"):
    gpu_id = int(os.getenv("LOCAL_RANK", 0)) if gpu_id is None else int(gpu_id)
    world_size = int(os.getenv("WORLD_SIZE", 1))
    try:
        import pynvml as nvml

        nvml.nvmlInit()
        if mode == "socket":
            cores = _gpu_cpu_list(nvml, gpu_id)
        elif mode == "single":
            cores = _gpu_cpu_list(nvml, gpu_id)[:1]
        elif mode == "single_unique":
            cores = (_socket_unique(nvml, gpu_id, world_size, True) or [])[:1]
        elif mode == "socket_unique_interleaved":
            cores = _socket_unique(nvml, gpu_id, world_size, True)
        elif mode == "socket_unique_continuous":
            cores = _socket_unique(nvml, gpu_id, world_size, False)
        else:
            raise RuntimeError("Unknown affinity mode")
        if cores:
            os.sched_setaffinity(0, cores)
    except RuntimeError:
        raise
    except Exception:  # noqa: BLE001 -- no NVML / no GPU / restricted cpuset: keep the inherited affinity
        pass
    return os.sched_getaffinity(0)
