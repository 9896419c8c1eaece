"""Host memory shared by the ranks of one box, page-locked in every process.

With several GPUs the value function and the policy of a sweep (24 MB for the 10^6 states of
config #5) used to reach the host through ONE GPU's PCIe link (`host_results = "root"`,
0.45 ms), or through all of them eight times over (`"all"`: eight full copies, 2.7 ms).  Here
the result arrays live in a POSIX shared-memory segment that every rank maps and registers
with its CUDA context (cudaHostRegister): each rank copies 1/N of the results over its own PCIe
link into the same pages, a flag barrier in the segment publishes them, and every rank reads
them in place.  The same segment carries the next call's input back: each rank uploads 1/N of
J and hands it to its peers over NVLink.

Layout of the segment:  control block | J staging | n_slots x (J [n_grid] | pol [n_grid][nc]).
Result slots go round a ring; a slot is reused only when no array handed out from it is alive
on ANY rank (liveness bits travel through the control block).  Nothing here computes: it is
plumbing for `DPSolver._sweep_host`.
"""
import mmap
import os
import time
import weakref

import numpy as np

__all__ = ["HostShare", "SlotArray"]

_CTRL_WORDS = 4096          # int64 words of the control block (16 per rank + header)


class _SlotToken(object):
    """alive as long as any array (or view of one) handed out from a result slot is"""
    __slots__ = ("__weakref__",)


class SlotArray(np.ndarray):
    """ndarray whose views carry the token of the result slot they look into (numpy collapses
    `.base` chains to the owner of the memory, so a plain view would not keep the array it was
    cut from - and with it the slot - alive); arrays computed from it own their data and carry
    nothing"""

    def __array_finalize__(self, obj):
        tok = getattr(obj, "_slot_token", None)
        self._slot_token = tok if (tok is not None and self.base is not None) else None

    def __array_wrap__(self, arr, context=None, return_scalar=False):
        # results of ufuncs / reductions are plain arrays (they do not look into the slot);
        # in-place operations (`J -= c`: the output IS a slot view) keep their token
        if isinstance(arr, SlotArray) and getattr(arr, "_slot_token", None) is not None:
            return arr
        arr = np.asarray(arr)
        if isinstance(arr, SlotArray):
            arr = arr.view(np.ndarray)
        return arr[()] if return_scalar else arr


class HostShare(object):
    def __init__(self, coll, n_grid, nc, cuda, n_slots=3):
        import torch
        self.coll = coll
        self.world, self.rank = coll.world, coll.rank
        self.n_grid, self.nc, self.n_slots = int(n_grid), int(nc), int(n_slots)
        self._cuda = bool(cuda)
        self._k = 0
        self._live = [0] * self.n_slots
        self._registered = False
        jb = 8 * self.n_grid
        pb = 8 * self.n_grid * max(self.nc, 1)
        self._off_stage = 8 * _CTRL_WORDS
        self._off_slot = [self._off_stage + jb + s * (jb + pb) for s in range(self.n_slots)]
        self.size = (self._off_slot[-1] + jb + pb + 4095) // 4096 * 4096
        name = None
        if self.rank == 0:
            # (a tmpfs smaller than the segment would only fail at the first page fault, with SIGBUS)
            st = os.statvfs("/dev/shm")
            if st.f_bavail * st.f_frsize < self.size + (64 << 20):
                name = ""
        if coll.all_gather_object(name)[0] == "":
            raise RuntimeError("/dev/shm is too small for %d MB of shared result slots" % (self.size >> 20))
        if self.rank == 0:
            name = "/dev/shm/sdp_b200_%d_%x" % (os.getpid(), int(time.time() * 1e6) & 0xffffffff)
            fd = os.open(name, os.O_CREAT | os.O_EXCL | os.O_RDWR, 0o600)
            os.ftruncate(fd, self.size)
        name = coll.all_gather_object(name)[0]
        if self.rank != 0:
            fd = os.open(name, os.O_RDWR)
        self.mm = mmap.mmap(fd, self.size)
        os.close(fd)
        coll.all_gather_object(True)          # everyone has mapped it ...
        if self.rank == 0:
            os.unlink(name)                   # ... so the name can go: nothing leaks on a crash
        self.buf = np.frombuffer(self.mm, dtype=np.uint8)
        self.ctrl = self.buf[:8 * _CTRL_WORDS].view(np.int64)
        if self.rank == 0:
            self.ctrl[:] = 0
        if self._cuda:
            rt = torch.cuda.cudart()
            err = rt.cudaHostRegister(self.buf.ctypes.data, self.size, 0)
            if int(err) != 0:
                raise RuntimeError("cudaHostRegister failed (%s)" % err)
            self._registered = True
        self.tensor = torch.from_numpy(self.buf)
        coll.all_gather_object(True)
        weakref.finalize(self, HostShare._release, self.buf.ctypes.data if self._registered else 0)

    @staticmethod
    def _release(addr):
        if addr:
            try:
                import torch
                torch.cuda.cudart().cudaHostUnregister(addr)
            except Exception:
                pass

    # -- views ----------------------------------------------------------------
    def _f64(self, off, n):
        return self.tensor[off:off + 8 * n].view(_torch_f64())

    def stage_J(self):
        return self._f64(self._off_stage, self.n_grid)

    def slot_J(self, s):
        return self._f64(self._off_slot[s], self.n_grid)

    def slot_pol(self, s):
        return self._f64(self._off_slot[s] + 8 * self.n_grid, self.n_grid * max(self.nc, 1))

    def slot_of(self, a):
        """index of the slot whose J block the host array `a` is (a view handed out earlier),
        -1 for the staging block, None for any other memory"""
        if not isinstance(a, np.ndarray) or a.dtype != np.float64 or a.size != self.n_grid \
                or not a.flags.c_contiguous:
            return None
        p = a.ctypes.data - self.buf.ctypes.data
        for s in range(self.n_slots):
            if p == self._off_slot[s]:
                return s
        return None

    # -- flag barrier / small broadcasts through the control block -------------
    def barrier(self, timeout=600.0):
        """every rank writes its sequence number, then waits for the others' (aligned 8-byte
        stores and loads on x86; the data they publish was written before the store)"""
        self._k += 1
        self.ctrl[16 * self.rank + 16] = self._k
        t0 = time.perf_counter()
        for r in range(self.world):
            spins = 0
            while self.ctrl[16 * r + 16] < self._k:
                spins += 1
                if spins > 2000:
                    time.sleep(0)
                    if time.perf_counter() - t0 > timeout:
                        raise RuntimeError("host barrier timed out: rank %d did not arrive" % r)

    def publish(self, word, value):
        self.ctrl[16 * self.rank + 17 + word] = int(value)

    def read(self, rank, word):
        return int(self.ctrl[16 * rank + 17 + word])

    # -- result slots ------------------------------------------------------------
    def live_mask(self):
        return sum(1 << s for s in range(self.n_slots) if self._live[s] > 0)

    def hand_out(self, s, shape_J, shape_pol, writable):
        """numpy views of slot `s` for the caller; the slot stays busy while they are alive"""
        n = self.n_grid
        J = self.buf[self._off_slot[s]:self._off_slot[s] + 8 * n].view(np.float64).reshape(shape_J)
        p0 = self._off_slot[s] + 8 * n
        pol = self.buf[p0:p0 + 8 * n * max(self.nc, 1)].view(np.float64)[:n * self.nc].reshape(shape_pol)
        tok = _SlotToken()
        self._live[s] += 1
        weakref.finalize(tok, self._drop, s)
        out = []
        for a in (J, pol):
            a = a.view(SlotArray)
            a._slot_token = tok
            a.flags.writeable = bool(writable)
            out.append(a)
        return out[0], out[1]

    def _drop(self, s):
        self._live[s] -= 1


def _torch_f64():
    import torch
    return torch.float64
