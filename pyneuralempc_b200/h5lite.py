"""A minimal, dependency-free reader for the HDF5 files Keras writes (``model.save("nn_model.h5")`` / ``save_weights``), enough to pull
the Dense kernels and biases of a ``Sequential`` model out of them.  The reference loads such a file through TensorFlow
(``examples/lotka_volterra/run.py:56``: ``tf.keras.models.load_model("./nn_model.h5")``) and hands the live model to ``KerasTFModel``
(``model/tensorflow.py:9-29``); the CUDA path only needs the numbers, and neither TensorFlow nor h5py is a dependency of this package.

Supported subset (what h5py's default ``libver='earliest'`` produces, which is what Keras uses): superblock version 0 or 1, version-1
object headers (with continuation blocks), old-style groups (symbol-table message -> v1 B-tree -> symbol-table nodes -> local heap),
contiguous or compact little-endian float datasets, fixed-length-string attributes stored inline (``layer_names`` / ``weight_names``).
Chunked / compressed datasets, new-style (fractal-heap) groups and variable-length data are rejected with ``ValueError``.

The activations are not stored with the weights but in the ``model_config`` JSON attribute (a variable-length string in the global heap);
``keras_activations`` finds the ``"activation": "<name>"`` entries of the Dense layers in that JSON by scanning the file bytes."""
from __future__ import annotations

import re
import struct

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class H5File:
    def __init__(self, path_or_bytes):
        self.buf = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray)) else open(path_or_bytes, "rb").read()
        b = self.buf
        if b[:8] != _SIG:
            raise ValueError("not an HDF5 file (signature missing at offset 0)")
        ver = b[8]
        if ver not in (0, 1):
            raise ValueError(f"HDF5 superblock version {ver} is not supported (only the version 0/1 files Keras / h5py write by default)")
        self.O, self.L = b[13], b[14]                               # size of offsets / lengths
        if self.O != 8 or self.L != 8:
            raise ValueError("only 8-byte offsets and lengths are supported")
        pos = 24 + (4 if ver == 1 else 0)                           # after group K values + consistency flags (+ v1: indexed storage K, reserved)
        self.base = self._u64(pos)
        root_ste = pos + 4 * 8                                      # base, free-space, end-of-file, driver-info addresses
        self.root = self._u64(root_ste + 8)                         # object header address of the root group

    # ---- primitives ------------------------------------------------------------------------------------------------------
    def _u16(self, p): return struct.unpack_from("<H", self.buf, p)[0]
    def _u32(self, p): return struct.unpack_from("<I", self.buf, p)[0]
    def _u64(self, p): return struct.unpack_from("<Q", self.buf, p)[0]

    def _messages(self, addr):
        """(type, flags, payload offset, payload size) of every message of a version-1 object header, continuation blocks included"""
        b = self.buf
        addr += self.base
        if b[addr] != 1:
            raise ValueError(f"object header version {b[addr]} at {addr} is not supported (new-style file?)")
        nmsg, hsize = self._u16(addr + 2), self._u32(addr + 8)
        blocks, out = [(addr + 16, hsize)], []
        while blocks and len(out) < nmsg:
            p, size = blocks.pop(0)
            end = p + size
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, flags = self._u16(p), self._u16(p + 2), b[p + 4]
                body = p + 8
                if mtype == 0x10:                                   # continuation: (offset, length)
                    blocks.append((self._u64(body) + self.base, self._u64(body + 8)))
                out.append((mtype, flags, body, msize))
                p = body + msize
        return out

    # ---- groups ----------------------------------------------------------------------------------------------------------
    def _heap_string(self, heap_addr, off):
        h = heap_addr + self.base
        if self.buf[h:h + 4] != b"HEAP":
            raise ValueError("local heap signature missing")
        data = self._u64(h + 24) + self.base
        end = self.buf.index(b"\0", data + off)
        return self.buf[data + off:end].decode("utf-8")

    def _btree_entries(self, node_addr, heap_addr, out):
        p = node_addr + self.base
        b = self.buf
        if b[p:p + 4] == b"SNOD":
            n = self._u16(p + 6)
            e = p + 8
            for _ in range(n):
                out[self._heap_string(heap_addr, self._u64(e))] = self._u64(e + 8)
                e += 40
            return
        if b[p:p + 4] != b"TREE" or b[p + 4] != 0:
            raise ValueError("expected a version-1 group B-tree node")
        n = self._u16(p + 6)
        q = p + 24 + 8                                              # skip key 0
        for _ in range(n):
            self._btree_entries(self._u64(q), heap_addr, out)
            q += 16                                                 # child address + next key

    def members(self, addr=None):
        """name -> object header address of the links of the group at `addr` (default: the root group), in B-tree (name) order"""
        addr = self.root if addr is None else addr
        for mtype, _, body, _ in self._messages(addr):
            if mtype == 0x11:                                       # symbol-table message: B-tree address, local-heap address
                out = {}
                self._btree_entries(self._u64(body), self._u64(body + 8), out)
                return out
        return None                                                 # not a group

    def walk(self, addr=None, prefix=""):
        """yields (path, object header address) of every dataset below the group"""
        mem = self.members(addr)
        for name, a in (mem or {}).items():
            if self.members(a) is None:
                yield prefix + name, a
            else:
                yield from self.walk(a, prefix + name + "/")

    def resolve(self, path):
        addr = self.root
        for part in [t for t in path.split("/") if t]:
            mem = self.members(addr)
            if mem is None or part not in mem:
                raise KeyError(path)
            addr = mem[part]
        return addr

    # ---- datasets and attributes ----------------------------------------------------------------------------------------------
    def _dataspace(self, body):
        b = self.buf
        ver, rank, flags = b[body], b[body + 1], b[body + 2]
        p = body + (8 if ver == 1 else 4)
        return tuple(self._u64(p + 8 * i) for i in range(rank))

    def _datatype(self, body):
        b = self.buf
        cls, bits0, size = b[body] & 0x0F, b[body + 1], self._u32(body + 4)
        if cls == 1:                                                # IEEE float
            if bits0 & 1:
                raise ValueError("big-endian floats are not supported")
            return np.dtype(f"<f{size}")
        if cls == 0:
            return np.dtype(f"<{'i' if b[body + 1] & 8 else 'u'}{size}")
        if cls == 3:                                                # fixed-length string
            return np.dtype(f"S{size}")
        raise ValueError(f"HDF5 datatype class {cls} is not supported")

    def dataset(self, addr):
        shape = dtype = None
        data = None
        for mtype, _, body, size in self._messages(addr):
            if mtype == 0x01:
                shape = self._dataspace(body)
            elif mtype == 0x03:
                dtype = self._datatype(body)
            elif mtype == 0x08:
                ver, cls = self.buf[body], self.buf[body + 1]
                if ver != 3:
                    raise ValueError(f"data layout message version {ver} is not supported")
                if cls == 1:
                    data = (self._u64(body + 2) + self.base, self._u64(body + 10))
                elif cls == 0:
                    data = (body + 4, self._u16(body + 2))
                else:
                    raise ValueError("chunked (compressed?) datasets are not supported: save the model without compression")
        if shape is None or dtype is None or data is None:
            raise ValueError("object is not a plain dataset")
        count = int(np.prod(shape)) if shape else 1
        if data[0] == _UNDEF + self.base or data[1] < count * dtype.itemsize:
            raise ValueError("dataset has no allocated storage")
        return np.frombuffer(self.buf, dtype, count, data[0]).reshape(shape).copy()

    def attributes(self, addr):
        """inline attributes (fixed-size types) of an object: name -> array; variable-length ones are skipped"""
        out = {}
        for mtype, _, body, _ in self._messages(addr):
            if mtype != 0x0C:
                continue
            b = self.buf
            ver = b[body]
            nsz, tsz, ssz = self._u16(body + 2), self._u16(body + 4), self._u16(body + 6)
            pad = (lambda v: (v + 7) & ~7) if ver == 1 else (lambda v: v)
            p = body + 8 + (1 if ver == 3 else 0)
            name = b[p:p + nsz].split(b"\0")[0].decode("utf-8")
            p += pad(nsz)
            try:
                dtype = self._datatype(p)
            except ValueError:
                continue
            shape = self._dataspace(p + pad(tsz))
            p += pad(tsz) + pad(ssz)
            count = int(np.prod(shape)) if shape else 1
            out[name] = np.frombuffer(b, dtype, count, p).reshape(shape).copy()
        return out


def keras_activations(buf):
    """activation names of the Dense layers, in model order, from the ``model_config`` JSON embedded in the file"""
    acts = []
    for m in re.finditer(rb'"class_name":\s*"Dense"', buf):
        a = re.compile(rb'"activation":\s*"(\w+)"').search(buf, m.end())
        if a:
            acts.append(a.group(1).decode())
    return acts


def read_keras_dense_stack(path_or_bytes):
    """-> (weights, activations): ``[(kernel[in, out], bias[out]), ...]`` in model order and the Dense activations (or None when the
    file carries no model config, e.g. ``save_weights``)."""
    f = H5File(path_or_bytes)
    top = f.members()
    base = top["model_weights"] if top and "model_weights" in top else f.root
    order = f.attributes(base).get("layer_names")
    groups = f.members(base)
    names = [n.decode() for n in order] if order is not None else list(groups)
    weights = []
    for lname in names:
        if lname not in groups:
            continue
        ds = {p.rsplit("/", 1)[-1]: f.dataset(a) for p, a in f.walk(groups[lname])}
        if not ds:
            continue                                                # InputLayer, Dropout, ...
        ker = [v for k, v in ds.items() if k.startswith("kernel") and v.ndim == 2]
        bia = [v for k, v in ds.items() if k.startswith("bias") and v.ndim == 1]
        if len(ker) != 1 or len(bia) != 1 or len(ds) != 2:
            raise ValueError(f"layer {lname!r} is not a Dense layer with a bias (datasets: {sorted(ds)})")
        weights.append((ker[0].astype(np.float64), bia[0].astype(np.float64)))
    if not weights:
        raise ValueError("no Dense layers found")
    acts = keras_activations(f.buf)
    return weights, (acts if len(acts) == len(weights) else None)
