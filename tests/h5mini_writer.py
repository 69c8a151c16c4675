"""A minimal HDF5 WRITER for tests: just enough of the file format (superblock version 0, version-1 object headers, old-style groups with
one symbol-table node each, contiguous little-endian datasets, fixed-length-string attributes) to produce the files Keras writes, so
that ``pyneuralempc_b200.h5lite`` and the ``.h5`` / ``.keras`` importers can be tested without h5py, TensorFlow or the reference's
fixture.  TEST INFRASTRUCTURE ONLY."""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class _Buf:
    def __init__(self):
        self.b = bytearray()

    def align(self):
        self.b += b"\0" * (-len(self.b) % 8)

    def put(self, data):
        self.align()
        addr = len(self.b)
        self.b += data
        return addr


def _msg(mtype, body):
    body = body + b"\0" * (-len(body) % 8)
    return struct.pack("<HHB3x", mtype, len(body), 0) + body


def _header(msgs):
    data = b"".join(msgs)
    return struct.pack("<BBHII4x", 1, 0, len(msgs), 1, len(data)) + data


def _dtype_msg(dt):
    dt = np.dtype(dt)
    if dt.kind == "f":
        props = {4: struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127), 8: struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)}[dt.itemsize]
        return struct.pack("<BBBBI", 0x11, 0x20, 0x1F if dt.itemsize == 4 else 0x3F, 0, dt.itemsize) + props
    if dt.kind in "ui":
        return struct.pack("<BBBBI", 0x10, 0x08 if dt.kind == "i" else 0, 0, 0, dt.itemsize) + struct.pack("<HH", 0, 8 * dt.itemsize)
    if dt.kind == "S":
        return struct.pack("<BBBBI", 0x13, 0, 0, 0, dt.itemsize)
    raise ValueError(dt)


def _space_msg(shape):
    return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", int(s)) for s in shape)


def _attr_msg(name, arr):
    arr = np.ascontiguousarray(arr)
    pad = lambda x: x + b"\0" * (-len(x) % 8)
    nm = name.encode() + b"\0"
    dt, sp = _dtype_msg(arr.dtype), _space_msg(arr.shape)
    return _msg(0x0C, struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(sp)) + pad(nm) + pad(dt) + pad(sp) + arr.tobytes())


def _write_dataset(buf, arr):
    arr = np.ascontiguousarray(arr)
    daddr = buf.put(arr.tobytes())
    msgs = [_msg(0x01, _space_msg(arr.shape)), _msg(0x03, _dtype_msg(arr.dtype)), _msg(0x08, struct.pack("<BBQQ", 3, 1, daddr, arr.nbytes))]
    return buf.put(_header(msgs))


def _write_group(buf, members, attrs=None):
    """members: dict name -> ndarray | (dict, attrs) | dict; returns the object-header address"""
    entries = []
    for name, val in members.items():
        if isinstance(val, tuple):
            addr = _write_group(buf, val[0], val[1])
        elif isinstance(val, dict):
            addr = _write_group(buf, val)
        else:
            addr = _write_dataset(buf, val)
        entries.append((name, addr))
    entries.sort()
    assert len(entries) <= 8, "one symbol-table node holds 8 links"
    heap_data = bytearray(b"\0" * 8)                       # offset 0: the empty name
    offs = []
    for name, _ in entries:
        offs.append(len(heap_data))
        heap_data += name.encode() + b"\0"
        heap_data += b"\0" * (-len(heap_data) % 8)
    hd = buf.put(bytes(heap_data))
    heap = buf.put(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), UNDEF, hd))
    snod = b"SNOD" + struct.pack("<BBH", 1, 0, len(entries))
    for (name, addr), off in zip(entries, offs):
        snod += struct.pack("<QQII16x", off, addr, 0, 0)
    snod += b"\0" * (40 * (8 - len(entries)))
    snod_addr = buf.put(snod)
    tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, UNDEF, UNDEF) + struct.pack("<QQQ", 0, snod_addr, offs[-1] if offs else 0)
    tree_addr = buf.put(tree)
    msgs = [_msg(0x11, struct.pack("<QQ", tree_addr, heap))]
    for k, v in (attrs or {}).items():
        msgs.append(_attr_msg(k, v))
    return buf.put(_header(msgs))


def write_h5(path, members, root_attrs=None):
    """members: nested dicts of numpy arrays; a group with attributes is given as ``(members, {attr_name: ndarray})``"""
    buf = _Buf()
    buf.b += b"\0" * 96                                    # superblock, filled in last
    root = _write_group(buf, members, root_attrs)
    buf.align()
    sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0) + struct.pack("<HHI", 4, 16, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, len(buf.b), UNDEF) + struct.pack("<QQII16x", 0, root, 0, 0)
    assert len(sb) == 96
    buf.b[:96] = sb
    with open(path, "wb") as f:
        f.write(bytes(buf.b))
    return path


def write_keras_h5(path, weights, activations, with_config=True):
    """a Keras ``model.save('x.h5')`` look-alike: model_weights/<layer>/<layer>/{kernel:0, bias:0}, the ``layer_names`` attribute and (as a
    byte dataset standing in for the variable-length ``model_config`` attribute) the model-config JSON"""
    import json
    names = ["dense" if i == 0 else f"dense_{i}" for i in range(len(weights))]
    mw = {}
    for nm, (W, b) in zip(names, weights):
        mw[nm] = {nm: {"kernel:0": np.asarray(W, np.float32), "bias:0": np.asarray(b, np.float32)}}
    S = max(len(n) for n in names)
    root = {"model_weights": (mw, {"layer_names": np.array([n.encode() for n in names], dtype=f"S{S}")})}
    if with_config:
        cfg = {"class_name": "Sequential", "config": {"layers": [{"class_name": "Dense", "config": {"name": n, "units": int(np.shape(W)[1]), "activation": a}}
                                                                   for n, (W, _), a in zip(names, weights, activations)]}}
        root["model_config_blob"] = np.frombuffer(json.dumps(cfg).encode(), dtype=np.uint8)
    return write_h5(path, root)


def write_keras_archive(path, weights, activations):
    """a Keras-3 ``.keras`` archive look-alike: config.json + model.weights.h5 (layers/<name>/vars/{0, 1})"""
    import json
    import os
    import tempfile
    import zipfile
    names = ["dense" if i == 0 else f"dense_{i}" for i in range(len(weights))]
    cfg = {"class_name": "Sequential", "config": {"name": "sequential", "layers": [{"class_name": "InputLayer", "config": {"name": "input_layer"}}] + [
        {"class_name": "Dense", "config": {"name": n, "units": int(np.shape(W)[1]), "activation": a}} for n, (W, _), a in zip(names, weights, activations)]}}
    layers = {n: {"vars": {"0": np.asarray(W, np.float32), "1": np.asarray(b, np.float32)}} for n, (W, b) in zip(names, weights)}
    with tempfile.TemporaryDirectory() as tmp:
        h5 = write_h5(os.path.join(tmp, "model.weights.h5"), {"layers": layers})
        with zipfile.ZipFile(path, "w") as zf:
            zf.writestr("config.json", json.dumps(cfg))
            zf.writestr("metadata.json", json.dumps({"keras_version": "3.0.0"}))
            zf.write(h5, "model.weights.h5")
    return path
