"""Device memory and zero-copy tensor exchange for libudal - no PyTorch involved.

``DeviceArray`` is a C-contiguous CUDA array owned by a ``Context``; it speaks both
``__cuda_array_interface__`` (v3) and DLPack (``__dlpack__`` / ``__dlpack_device__``), so any
framework can consume it without a copy.  ``as_device`` accepts NumPy arrays (copied to the
device), ``DeviceArray``s and any foreign object exposing ``__cuda_array_interface__`` or
``__dlpack__`` (borrowed, zero copy).
"""
import ctypes

import numpy as np

from . import _lib

# ---- minimal DLPack ABI (dlpack.h v0.8) ----------------------------------------------------------
kDLCUDA = 2
kDLCUDAHost = 3
_DL_CODES = {"i": 0, "u": 1, "f": 2}
_DL_KINDS = {0: "i", 1: "u", 2: "f"}


class DLDevice(ctypes.Structure):
    _fields_ = [("device_type", ctypes.c_int), ("device_id", ctypes.c_int)]


class DLDataType(ctypes.Structure):
    _fields_ = [("code", ctypes.c_uint8), ("bits", ctypes.c_uint8), ("lanes", ctypes.c_uint16)]


class DLTensor(ctypes.Structure):
    _fields_ = [("data", ctypes.c_void_p), ("device", DLDevice), ("ndim", ctypes.c_int),
                ("dtype", DLDataType), ("shape", ctypes.POINTER(ctypes.c_int64)),
                ("strides", ctypes.POINTER(ctypes.c_int64)), ("byte_offset", ctypes.c_uint64)]


class DLManagedTensor(ctypes.Structure):
    pass


_DELETER = ctypes.CFUNCTYPE(None, ctypes.POINTER(DLManagedTensor))
DLManagedTensor._fields_ = [("dl_tensor", DLTensor), ("manager_ctx", ctypes.c_void_p),
                            ("deleter", _DELETER)]

_py = ctypes.pythonapi
_py.PyCapsule_New.restype = ctypes.py_object
_py.PyCapsule_New.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p]
_py.PyCapsule_GetPointer.restype = ctypes.c_void_p
_py.PyCapsule_GetPointer.argtypes = [ctypes.py_object, ctypes.c_char_p]
_py.PyCapsule_IsValid.restype = ctypes.c_int
_py.PyCapsule_IsValid.argtypes = [ctypes.py_object, ctypes.c_char_p]
_py.PyCapsule_SetName.restype = ctypes.c_int
_py.PyCapsule_SetName.argtypes = [ctypes.py_object, ctypes.c_char_p]

_exported = {}  # id -> (managed tensor, shape array, owner) kept alive until the consumer's deleter runs


@_DELETER
def _export_deleter(ptr):
    _exported.pop(ctypes.addressof(ptr.contents), None)


class Context:
    """One libudal context = one GPU + one stream (include/udal.h udal_ctx)."""

    def __init__(self, cfg):
        self.lib = _lib.load()
        self.cfg = cfg
        h = ctypes.c_void_p()
        _lib.check(self.lib.udal_create(ctypes.byref(cfg), ctypes.byref(h)))
        self.handle = h
        self.device = cfg.device
        self._pool = {}  # nbytes -> [device pointers]: stream-ordered reuse, no cudaMalloc per call

    def close(self):
        if getattr(self, "handle", None):
            self._pool.clear()
            self.lib.udal_destroy(self.handle)  # frees every allocation of the context
            self.handle = None

    def _alloc(self, nbytes):
        free = self._pool.get(nbytes)
        if free:
            return free.pop()
        p = ctypes.c_void_p()
        _lib.check(self.lib.udal_malloc(self.handle, nbytes, ctypes.byref(p)))
        return p.value

    def _release(self, ptr, nbytes):
        # every entry point orders its work behind everything the context enqueued before (udal_join brings the
        # post stream of udal_run back in), so a released block can be handed out again without a sync
        self._pool.setdefault(nbytes, []).append(ptr)

    def trim(self):
        """Return pooled blocks to the driver."""
        for ptrs in self._pool.values():
            for p in ptrs:
                self.lib.udal_free(self.handle, ctypes.c_void_p(p))
        self._pool.clear()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- memory --------------------------------------------------------------------------------
    def empty(self, shape, dtype=np.float32):
        return DeviceArray(self, shape, dtype)

    def zeros(self, shape, dtype=np.float32):
        a = DeviceArray(self, shape, dtype)
        _lib.check(self.lib.udal_memset(self.handle, a.ptr, 0, a.nbytes))
        return a

    def to_device(self, host):
        host = np.ascontiguousarray(host)
        a = DeviceArray(self, host.shape, host.dtype)
        a.copy_from_host(host)
        return a

    def set_stream(self, cuda_stream_handle):
        _lib.check(self.lib.udal_set_stream(self.handle, ctypes.c_void_p(cuda_stream_handle or 0)))

    def sync(self):
        _lib.check(self.lib.udal_sync(self.handle))

    def stream_handle(self):
        p = ctypes.c_void_p()
        _lib.check(self.lib.udal_get_stream(self.handle, ctypes.byref(p)))
        return p.value or 0

    def wait_stream(self, producer_stream):
        """Order the context's later work after everything enqueued so far on ``producer_stream`` (int handle;
        0 / 1 = legacy default stream, 2 = per-thread default stream)."""
        _lib.check(self.lib.udal_wait_stream(self.handle, ctypes.c_void_p(producer_stream or 0)))

    def wait_context(self, producer):
        """orders all later work of this context behind everything ``producer`` (another Context) has enqueued"""
        if producer is not self and producer.handle:
            _lib.check(self.lib.udal_wait_context(self.handle, producer.handle))

    def timer_start(self):
        _lib.check(self.lib.udal_timer_start(self.handle))

    def timer_stop(self):
        ms = ctypes.c_float(0)
        _lib.check(self.lib.udal_timer_stop(self.handle, ctypes.byref(ms)))
        return ms.value

    def launch_count(self):
        n = ctypes.c_int64(0)
        _lib.check(self.lib.udal_launch_count(self.handle, ctypes.byref(n)))
        return n.value

    def layer_times(self, fn):
        """CUDA-event time of every head-tower layer launched by ``fn()`` (udal_profile_layers):
        list of ms, class tower layers 0..R then box tower layers 0..R."""
        _lib.check(self.lib.udal_profile_layers(self.handle, 1))
        try:
            fn()
            ms = (ctypes.c_float * 64)()
            n = ctypes.c_int(0)
            _lib.check(self.lib.udal_get_layer_times(self.handle, ms, 64, ctypes.byref(n)))
        finally:
            _lib.check(self.lib.udal_profile_layers(self.handle, 0))
        return [float(ms[i]) for i in range(n.value)]

    def scratch_bytes(self):
        n = ctypes.c_size_t(0)
        _lib.check(self.lib.udal_scratch_bytes(self.handle, ctypes.byref(n)))
        return n.value


class PinnedArray:
    """Page-locked host staging buffer viewed as a NumPy array (``.array``)."""

    def __init__(self, shape, dtype=np.float32):
        self.lib = _lib.load()
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize
        p = ctypes.c_void_p()
        _lib.check(self.lib.udal_host_alloc(self.nbytes, ctypes.byref(p)))
        self.ptr = p.value
        buf = (ctypes.c_char * max(self.nbytes, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape, dtype=np.int64))).reshape(self.shape)

    def __del__(self):
        try:
            if self.ptr:
                self.array = None
                self.lib.udal_host_free(ctypes.c_void_p(self.ptr))
                self.ptr = None
        except Exception:
            pass


class DeviceArray:
    """C-contiguous array in the memory of a Context's GPU (owned, or borrowed from a foreign
    producer when ``base`` is given)."""

    def __init__(self, ctx, shape, dtype=np.float32, ptr=None, base=None):
        self.ctx = ctx
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        self.size = int(np.prod(self.shape, dtype=np.int64))
        self.nbytes = self.size * self.dtype.itemsize
        self._owned = ptr is None
        self._base = base
        if ptr is None:
            ptr = ctx._alloc(self.nbytes)
        self.ptr = ptr

    def __del__(self):
        try:
            if self._owned and self.ptr and self.ctx.handle:
                self.ctx._release(self.ptr, self.nbytes)
            self.ptr = None
        except Exception:
            pass

    @property
    def ndim(self):
        return len(self.shape)

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        shape = list(shape)
        if -1 in shape:
            i = shape.index(-1)
            known = int(np.prod([s for s in shape if s != -1], dtype=np.int64))
            shape[i] = self.size // max(known, 1)
        assert int(np.prod(shape, dtype=np.int64)) == self.size
        return DeviceArray(self.ctx, shape, self.dtype, ptr=self.ptr, base=self)

    def slice0(self, start, stop):
        """View of rows [start, stop) along axis 0."""
        row = (self.size // self.shape[0]) * self.dtype.itemsize if self.shape[0] else 0
        return DeviceArray(self.ctx, (stop - start,) + self.shape[1:], self.dtype,
                           ptr=self.ptr + start * row, base=self)

    def copy_from_host(self, host):
        host = np.ascontiguousarray(host, dtype=self.dtype)
        assert host.size == self.size, (host.shape, self.shape)
        _lib.check(self.ctx.lib.udal_memcpy_h2d(self.ctx.handle, self.ptr, host.ctypes.data, self.nbytes))
        # pageable sources are staged synchronously by the driver; pinned ones are async:
        self._keep = host
        return self

    def copy_to_host(self, out=None, sync=True):
        if out is None:
            out = np.empty(self.shape, self.dtype)
        assert out.flags.c_contiguous and out.nbytes == self.nbytes
        _lib.check(self.ctx.lib.udal_memcpy_d2h(self.ctx.handle, out.ctypes.data, self.ptr, self.nbytes))
        if sync:
            self.ctx.sync()
        return out

    def numpy(self):
        return self.copy_to_host()

    def __array__(self, dtype=None, copy=None):
        a = self.copy_to_host()
        return a if dtype is None else a.astype(dtype)

    # -- zero-copy export ----------------------------------------------------------------------
    @property
    def __cuda_array_interface__(self):
        # "stream": None tells the consumer that no synchronisation is needed - so make that true: results may
        # still be pending on the context's streams (udal_run's tail runs on the post stream)
        self.ctx.sync()
        return {"shape": self.shape, "typestr": self.dtype.str, "data": (self.ptr or 0, False),
                "version": 3, "strides": None, "stream": None}

    def __dlpack_device__(self):
        return (kDLCUDA, self.ctx.device)

    def __dlpack__(self, stream=None, **_):
        self.ctx.sync()  # the consumer may use any stream
        m = DLManagedTensor()
        shape = (ctypes.c_int64 * max(self.ndim, 1))(*self.shape)
        m.dl_tensor.data = self.ptr
        m.dl_tensor.device = DLDevice(kDLCUDA, self.ctx.device)
        m.dl_tensor.ndim = self.ndim
        m.dl_tensor.dtype = DLDataType(_DL_CODES[self.dtype.kind], self.dtype.itemsize * 8, 1)
        m.dl_tensor.shape = ctypes.cast(shape, ctypes.POINTER(ctypes.c_int64))
        m.dl_tensor.strides = None
        m.dl_tensor.byte_offset = 0
        m.manager_ctx = None
        m.deleter = _export_deleter
        _exported[ctypes.addressof(m)] = (m, shape, self)
        return _py.PyCapsule_New(ctypes.addressof(m), b"dltensor", None)


def _from_dlpack(ctx, obj):
    # DLPack protocol: the consumer passes ITS stream; the producer makes that stream wait for the work that
    # writes the tensor.  The context's streams are non-blocking, i.e. they do not implicitly follow the
    # producer's default stream.  (A CUDA stream handle is never 0 here; 1 / 2 are the reserved default handles.)
    try:
        cap = obj.__dlpack__(stream=ctx.stream_handle())
    except (TypeError, ValueError, AssertionError):
        cap = obj.__dlpack__()
        ctx.wait_stream(0)  # producer did not take a stream: order behind the legacy default stream
    if not _py.PyCapsule_IsValid(cap, b"dltensor"):
        raise ValueError("object did not return a 'dltensor' capsule")
    mptr = _py.PyCapsule_GetPointer(cap, b"dltensor")
    managed = ctypes.cast(mptr, ctypes.POINTER(DLManagedTensor))
    t = managed.contents.dl_tensor
    if t.device.device_type != kDLCUDA or t.device.device_id != ctx.device:
        raise ValueError("DLPack tensor is not on cuda:%d" % ctx.device)
    if t.dtype.lanes != 1:
        raise ValueError("vector dtypes are not supported")
    shape = tuple(t.shape[i] for i in range(t.ndim))
    dtype = np.dtype("%s%d" % (_DL_KINDS[t.dtype.code], t.dtype.bits // 8))
    if t.strides:
        expect = 1
        for i in range(t.ndim - 1, -1, -1):
            if shape[i] != 1 and t.strides[i] != expect:
                raise ValueError("DLPack tensor must be C-contiguous")
            expect *= shape[i]
    _py.PyCapsule_SetName(cap, b"used_dltensor")

    class _Owner:
        def __del__(self_inner):
            if managed.contents.deleter:
                managed.contents.deleter(managed)

    owner = _Owner()
    owner.capsule = cap
    return DeviceArray(ctx, shape, dtype, ptr=(t.data or 0) + t.byte_offset, base=owner)


def is_device_array(x):
    """DeviceArray or a foreign CUDA-array-interface producer.  Never touches a DeviceArray's own
    ``__cuda_array_interface__``: that property synchronises the context before it answers."""
    return isinstance(x, DeviceArray) or hasattr(x, "__cuda_array_interface__")


def as_device(ctx, x, dtype=None):
    """-> (DeviceArray, was_host).  Host arrays are copied, device arrays are borrowed."""
    if isinstance(x, DeviceArray):
        if dtype is not None and x.dtype != np.dtype(dtype):
            raise TypeError("expected %s, got %s" % (np.dtype(dtype), x.dtype))
        if x.ctx is not ctx:   # produced by another context (own streams): order this context behind it
            ctx.wait_context(x.ctx)
        return x, False
    cai = getattr(x, "__cuda_array_interface__", None)
    if cai is not None:
        if cai.get("strides") is not None:
            shape, strides = cai["shape"], cai["strides"]
            expect = np.dtype(cai["typestr"]).itemsize
            for s, st in zip(reversed(shape), reversed(strides)):
                if s != 1 and st != expect:
                    raise ValueError("device array must be C-contiguous")
                expect *= s
        arr = DeviceArray(ctx, cai["shape"], np.dtype(cai["typestr"]), ptr=cai["data"][0], base=x)
        # stream ordering (CAI v3): an int names the producer's stream; None / absent (v2 producers such as torch) says
        # nothing - order behind the legacy default stream, where such producers run unless told otherwise
        ctx.wait_stream(int(cai.get("stream") or 0))
    elif hasattr(x, "__dlpack__") and not isinstance(x, np.ndarray):
        arr = _from_dlpack(ctx, x)
    else:
        host = np.asarray(x)
        if dtype is not None:
            host = host.astype(dtype, copy=False)
        return ctx.to_device(host), True
    if dtype is not None and arr.dtype != np.dtype(dtype):
        raise TypeError("expected %s, got %s" % (np.dtype(dtype), arr.dtype))
    return arr, False


def ptr_array(arrays):
    """list of DeviceArray -> ctypes void*[MAX_LEVELS]"""
    out = (ctypes.c_void_p * _lib.MAX_LEVELS)()
    for i, a in enumerate(arrays):
        out[i] = a.ptr
    return out
