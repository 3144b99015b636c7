"""Reader / writer for the Torch7 (2012) binary serialisation the reference stores everything in:
`torch.save` / `torch.load` of calibration files (radial/generate_calibration_file.lua:106-114,
radial/*.cal), trained models (opticalflow_model_io.lua:98-207, version 9;
radial/radial_opticalflow_network.lua:120-158, version 1), cached ground-truth flows
(radial/radial_opticalflow_data.lua:41) and score tables.

Format (torch7 File.lua, binary DiskFile, little endian, as found in radial/gopro.cal):
  object  := int32 type, payload
  type 0 nil | 1 number: float64 | 2 string: int32 n, n bytes | 5 boolean: int32
  type 3 table: int32 index, [int32 size, size x (key object, value object)]   (body once per index)
  type 4 torch object: int32 index, [string "V 1", string class, class payload]
  type 6 function: int32 index, [int32 n, n bytes of string.dump, upvalue object]
  Tensor payload : int32 nDim, nDim x int64 size, nDim x int64 stride, int64 offset (1-based),
                   storage object (or nil)
  Storage payload: int64 n, n raw elements
Tables with keys 1..n come back as lists, other tables as dicts (insertion order = file order);
tensors as numpy arrays (views keep their strides); functions and unknown classes as opaque
records that write back byte for byte.  Function records are restated from File.lua without a
fixture in the reference tree (no saved model is shipped): parsed, never executed.
"""
import struct

import numpy as np

TYPE_NIL, TYPE_NUMBER, TYPE_STRING, TYPE_TABLE, TYPE_TORCH, TYPE_BOOLEAN, TYPE_FUNCTION = 0, 1, 2, 3, 4, 5, 6
TYPE_LEGACY_RECUR_FUNCTION, TYPE_RECUR_FUNCTION = 7, 8

_STORAGES = {
    "torch.FloatStorage": np.float32, "torch.DoubleStorage": np.float64, "torch.LongStorage": np.int64,
    "torch.IntStorage": np.int32, "torch.ShortStorage": np.int16, "torch.ByteStorage": np.uint8,
    "torch.CharStorage": np.int8,
}
_TENSORS = {k.replace("Storage", "Tensor"): v for k, v in _STORAGES.items()}
_TENSOR_OF = {np.dtype(v): k for k, v in _TENSORS.items()}


class Torch7FormatError(ValueError):
    pass


class LuaFunction:
    """string.dump bytecode + upvalues; opaque."""

    def __init__(self, dumped, upvalues):
        self.dumped, self.upvalues = dumped, upvalues

    def __repr__(self):
        return "<lua function, %d bytes>" % len(self.dumped)


class TorchObject:
    """A torch class this reader has no decoder for (nn modules ...): class name + the table its
    default write() stores."""

    def __init__(self, typename, fields, version=1):
        self.typename, self.fields, self.version = typename, fields, version

    def __repr__(self):
        return "<%s>" % self.typename


class Table(dict):
    """A Lua table with non-sequence keys.  Attribute access like the Lua code (t.wImg)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k) from None


class _Reader:
    def __init__(self, data):
        self.d, self.p, self.objects = data, 0, {}

    def take(self, fmt):
        n = struct.calcsize(fmt)
        if self.p + n > len(self.d):
            raise Torch7FormatError("truncated file at byte %d" % self.p)
        v = struct.unpack_from(fmt, self.d, self.p)
        self.p += n
        return v[0] if len(v) == 1 else v

    def raw(self, n):
        if n < 0 or self.p + n > len(self.d):
            raise Torch7FormatError("truncated file at byte %d (wanted %d bytes)" % (self.p, n))
        b = self.d[self.p:self.p + n]
        self.p += n
        return b

    def string(self):
        return self.raw(self.take("<i")).decode("latin-1")

    def obj(self):
        t = self.take("<i")
        if t == TYPE_NIL:
            return None
        if t == TYPE_NUMBER:
            v = self.take("<d")
            return int(v) if v == int(v) and abs(v) < 2 ** 53 else v
        if t == TYPE_BOOLEAN:
            return self.take("<i") == 1
        if t == TYPE_STRING:
            return self.string()
        if t in (TYPE_TABLE, TYPE_TORCH, TYPE_FUNCTION, TYPE_RECUR_FUNCTION, TYPE_LEGACY_RECUR_FUNCTION):
            index = self.take("<i")
            if index in self.objects:
                return self.objects[index]
            if t == TYPE_TABLE:
                return self.table(index)
            if t == TYPE_TORCH:
                return self.torch(index)
            n = self.take("<i")
            f = LuaFunction(self.raw(n), None)
            self.objects[index] = f
            f.upvalues = self.obj()
            return f
        raise Torch7FormatError("unknown type tag %d at byte %d (an ascii-mode file?)" % (t, self.p - 4))

    def table(self, index):
        n = self.take("<i")
        t = Table()
        self.objects[index] = t
        for _ in range(n):
            k = self.obj()
            t[k] = self.obj()
        if n and all(isinstance(k, int) for k in t) and sorted(t) == list(range(1, n + 1)):
            seq = [t[i] for i in range(1, n + 1)]
            self.objects[index] = seq   # references read later resolve to the list
            return seq
        return t

    def torch(self, index):
        version = self.string()
        if version.startswith("V "):
            vnum, cls = int(version[2:]), self.string()
        else:                        # pre-versioning files: the string is the class name
            vnum, cls = 0, version
        if cls in _STORAGES:
            n = self.take("<q")
            dt = np.dtype(_STORAGES[cls])
            a = np.frombuffer(self.raw(n * dt.itemsize), dtype=dt).copy()
            self.objects[index] = a
            return a
        if cls in _TENSORS:
            nd = self.take("<i")
            sizes = [self.take("<q") for _ in range(nd)]
            strides = [self.take("<q") for _ in range(nd)]
            offset = self.take("<q") - 1
            placeholder = object()
            self.objects[index] = placeholder
            storage = self.obj()
            dt = np.dtype(_TENSORS[cls])
            if storage is None or nd == 0:
                a = np.zeros((0,), dt)
            else:
                need = offset + sum((s - 1) * st for s, st in zip(sizes, strides)) + 1
                if offset < 0 or need > storage.size or any(s < 0 for s in sizes):
                    raise Torch7FormatError("%s of size %s does not fit its storage of %d" % (cls, sizes, storage.size))
                a = np.lib.stride_tricks.as_strided(storage[offset:], shape=sizes,
                                                    strides=[st * dt.itemsize for st in strides])
            self.objects[index] = a
            return a
        o = TorchObject(cls, None, vnum)
        self.objects[index] = o
        o.fields = self.obj()
        return o


def loads(data):
    r = _Reader(bytes(data))
    v = r.obj()
    if r.p != len(r.d):
        raise Torch7FormatError("%d trailing bytes" % (len(r.d) - r.p))
    return v


def load(path):
    with open(path, "rb") as f:
        return loads(f.read())


class _Writer:
    def __init__(self):
        self.out, self.ids, self.next, self.keep = [], {}, 1, []

    def put(self, fmt, *v):
        self.out.append(struct.pack(fmt, *v))

    def string(self, s):
        b = s.encode("latin-1")
        self.put("<i", len(b))
        self.out.append(b)

    def ref(self, tag, o):
        """Writes the tag + index; True when the body is still to be written."""
        self.put("<i", tag)
        known = id(o) in self.ids
        if not known:
            self.ids[id(o)] = self.next
            self.keep.append(o)
            self.next += 1
        self.put("<i", self.ids[id(o)])
        return not known

    def obj(self, o):
        if o is None:
            self.put("<i", TYPE_NIL)
        elif isinstance(o, (bool, np.bool_)):
            self.put("<ii", TYPE_BOOLEAN, 1 if o else 0)
        elif isinstance(o, (int, float, np.integer, np.floating)):
            self.put("<id", TYPE_NUMBER, float(o))
        elif isinstance(o, str):
            self.put("<i", TYPE_STRING)
            self.string(o)
        elif isinstance(o, dict):
            if self.ref(TYPE_TABLE, o):
                self.put("<i", len(o))
                for k, v in o.items():
                    self.obj(k)
                    self.obj(v)
        elif isinstance(o, (list, tuple)):
            if self.ref(TYPE_TABLE, o):
                self.put("<i", len(o))
                for i, v in enumerate(o):
                    self.obj(i + 1)
                    self.obj(v)
        elif isinstance(o, np.ndarray):
            self.tensor(o)
        elif isinstance(o, LuaFunction):
            if self.ref(TYPE_FUNCTION, o):
                self.put("<i", len(o.dumped))
                self.out.append(o.dumped)
                self.obj(o.upvalues)
        elif isinstance(o, TorchObject):
            if self.ref(TYPE_TORCH, o):
                self.string("V %d" % o.version)
                self.string(o.typename)
                self.obj(o.fields)
        else:
            raise Torch7FormatError("cannot serialise %r" % type(o))

    def tensor(self, a):
        if a.dtype not in _TENSOR_OF:
            raise Torch7FormatError("no torch tensor type for dtype %s" % a.dtype)
        cls = _TENSOR_OF[a.dtype]
        if not self.ref(TYPE_TORCH, a):
            return
        self.string("V 1")
        self.string(cls)
        # like torch.save of a freshly made tensor: its own contiguous storage, offset 1
        c = np.ascontiguousarray(a)
        self.put("<i", c.ndim)
        for s in c.shape:
            self.put("<q", s)
        for st in c.strides:
            self.put("<q", st // c.itemsize)
        self.put("<q", 1)
        storage = c.reshape(-1)
        self.ref(TYPE_TORCH, storage)
        self.string("V 1")
        self.string(cls.replace("Tensor", "Storage"))
        self.put("<q", storage.size)
        self.out.append(storage.tobytes())


def dumps(obj):
    w = _Writer()
    w.obj(obj)
    return b"".join(w.out)


def save(path, obj):
    with open(path, "wb") as f:
        f.write(dumps(obj))
