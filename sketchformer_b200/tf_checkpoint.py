"""TensorFlow tensor-bundle checkpoints without TensorFlow.

The reference saves `tf.train.Checkpoint(transformer=self, optimizer=self.optimizer)` through a CheckpointManager
(core/models.py:321-331): a pair of files  <prefix>.index  (a LevelDB-format sorted string table that maps
object-graph keys such as `transformer/encoder/enc_layers/0/mha/wq/kernel/.ATTRIBUTES/VARIABLE_VALUE` to
BundleEntryProto records) and  <prefix>.data-00000-of-00001  (the raw little-endian tensor bytes).  This module
reads and writes that pair and maps the keys to this package's parameter names (init.py), so the published
pretrained models (reference README.md:162-164) load into the arena, and weights trained here can be handed back.

Format restated from the published TensorFlow / LevelDB specifications (tensor_bundle.proto, table_format.md):
  table file  = data blocks | metaindex block | index block | 48-byte footer (two block handles, magic)
  block       = prefix-compressed entries, restart array, restart count ; trailer = type byte + masked crc32c
  entry       = varint shared, varint non_shared, varint value_len, key suffix, value
  key ""      -> BundleHeaderProto{num_shards, endianness, version}; other keys -> BundleEntryProto{dtype, shape,
                 shard_id, offset, size, crc32c}
There is no TensorFlow in this image and no reference checkpoint on disk, so the reader is checked against the
writer below and against hand-assembled blocks (tests/test_tf_checkpoint.py), not against a file TF wrote:
"parity unpinned" like the oracle.  The one string tensor of such a bundle, `_CHECKPOINTABLE_OBJECT_GRAPH` (a
serialized TrackableObjectGraph: the tree of attribute names `tf.train.Checkpoint.restore` walks to find each
variable's key, plus the optimizer's slot-variable table), is written by export_checkpoint() and parsed by
read_object_graph(); the numeric reader does not need it (keys are looked up by name)."""
import os
import struct
from typing import Dict, Iterable, List, Tuple

import numpy as np

TABLE_MAGIC = 0xDB4775248B80FB57
SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"
# tensorflow/core/framework/types.proto
DTYPES = {1: np.dtype("<f4"), 2: np.dtype("<f8"), 3: np.dtype("<i4"), 4: np.dtype("u1"), 5: np.dtype("<i2"),
          6: np.dtype("i1"), 9: np.dtype("<i8"), 10: np.dtype("?"), 17: np.dtype("<u2"), 22: np.dtype("<u4"),
          23: np.dtype("<u8")}
DT_STRING = 7
DTYPE_IDS = {v: k for k, v in DTYPES.items()}


# ---- crc32c (Castagnoli), masked the LevelDB way -----------------------------------------------------------
def _make_crc_table():
    t = np.zeros(256, dtype=np.uint32)
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ (0x82F63B78 if c & 1 else 0)
        t[i] = c
    return [int(x) for x in t]


_CRC = _make_crc_table()


def crc32c(data: bytes, crc: int = 0) -> int:
    c = crc ^ 0xFFFFFFFF
    tab = _CRC
    for b in data:
        c = tab[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def mask_crc(crc: int) -> int:
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


# ---- varints / minimal protobuf --------------------------------------------------------------------------
def _get_varint(buf, pos):
    out = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if b < 0x80:
            return out, pos
        shift += 7


def _put_varint(v: int) -> bytes:
    if v < 0:
        v += 1 << 64
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _pb_fields(buf) -> Iterable[Tuple[int, int, object]]:
    """(field number, wire type, value) of one serialized message; value is int or bytes"""
    pos, n = 0, len(buf)
    while pos < n:
        tag, pos = _get_varint(buf, pos)
        f, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            ln, pos = _get_varint(buf, pos)
            v = bytes(buf[pos:pos + ln])
            pos += ln
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError(f"protobuf wire type {wt} not supported")
        yield f, wt, v


def _signed64(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def _parse_entry(buf):
    e = {"dtype": 0, "shape": [], "shard_id": 0, "offset": 0, "size": 0, "crc32c": None, "sliced": False}
    for f, _, v in _pb_fields(buf):
        if f == 1:
            e["dtype"] = v
        elif f == 2:                       # TensorShapeProto: repeated Dim dim = 2 { int64 size = 1 }
            for f2, _, v2 in _pb_fields(v):
                if f2 == 2:
                    size = 0
                    for f3, _, v3 in _pb_fields(v2):
                        if f3 == 1:
                            size = _signed64(v3)
                    e["shape"].append(size)
        elif f == 3:
            e["shard_id"] = v
        elif f == 4:
            e["offset"] = v
        elif f == 5:
            e["size"] = v
        elif f == 6:
            e["crc32c"] = v
        elif f == 7:
            e["sliced"] = True
    return e


def _entry_proto(dtype_id, shape, offset, size, crc) -> bytes:
    dims = b"".join(b"\x12" + _put_varint(len(d)) + d for d in (b"\x08" + _put_varint(int(s)) for s in shape))
    out = b"\x08" + _put_varint(dtype_id)
    out += b"\x12" + _put_varint(len(dims)) + dims
    if offset:
        out += b"\x20" + _put_varint(offset)
    out += b"\x28" + _put_varint(size)
    out += b"\x35" + struct.pack("<I", crc)
    return out


# ---- snappy raw format (LevelDB blocks may be compressed; TF's BundleWriter normally stores them raw) ----------
def snappy_uncompress(buf: bytes) -> bytes:
    n, pos = _get_varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(buf[pos:pos + nb], "little")
                pos += nb
            ln += 1
            out += buf[pos:pos + ln]
            pos += ln
            continue
        if kind == 1:
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | buf[pos]
            pos += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = buf[pos] | (buf[pos + 1] << 8)
            pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 4], "little")
            pos += 4
        if off == 0 or off > len(out):
            raise ValueError("snappy: bad copy offset")
        for _ in range(ln):                 # copies may overlap their own output
            out.append(out[-off])
    if len(out) != n:
        raise ValueError("snappy: length mismatch")
    return bytes(out)


# ---- sorted string table -----------------------------------------------------------------------------------
def _read_block(data: bytes, off: int, size: int, verify: bool) -> bytes:
    raw = data[off:off + size]
    ctype = data[off + size]
    if verify:
        want = struct.unpack_from("<I", data, off + size + 1)[0]
        got = mask_crc(crc32c(data[off:off + size + 1]))
        if want != got:
            raise ValueError(f"table block at {off}: crc mismatch")
    if ctype == 0:
        return raw
    if ctype == 1:
        return snappy_uncompress(raw)
    raise ValueError(f"table block at {off}: unknown compression type {ctype}")


def _block_entries(block: bytes) -> List[Tuple[bytes, bytes]]:
    nrestart = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * nrestart
    pos, key, out = 0, b"", []
    while pos < end:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        out.append((key, block[pos:pos + vlen]))
        pos += vlen
    return out


def read_table(path: str, verify: bool = True) -> List[Tuple[bytes, bytes]]:
    with open(path, "rb") as f:
        data = f.read()
    if len(data) < 48 or struct.unpack_from("<Q", data, len(data) - 8)[0] != TABLE_MAGIC:
        raise ValueError(f"{path}: not a sorted string table (bad magic)")
    foot = data[-48:]
    _, p = _get_varint(foot, 0)            # metaindex handle (unused)
    _, p = _get_varint(foot, p)
    ioff, p = _get_varint(foot, p)
    isize, p = _get_varint(foot, p)
    out = []
    for _, handle in _block_entries(_read_block(data, ioff, isize, verify)):
        boff, q = _get_varint(handle, 0)
        bsize, _ = _get_varint(handle, q)
        out.extend(_block_entries(_read_block(data, boff, bsize, verify)))
    return out


def _build_block(items: List[Tuple[bytes, bytes]], restart_interval: int = 16) -> bytes:
    out, restarts, prev = bytearray(), [], b""
    for i, (k, v) in enumerate(items):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            m = min(len(prev), len(k))
            while shared < m and prev[shared] == k[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v)) + k[shared:] + v
        prev = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def write_table(path: str, items: List[Tuple[bytes, bytes]], block_size: int = 4096):
    items = sorted(items)
    out = bytearray()

    def emit(block: bytes):
        off = len(out)
        out.extend(block)
        out.append(0)                                                   # kNoCompression
        out.extend(struct.pack("<I", mask_crc(crc32c(block + b"\x00"))))
        return _put_varint(off) + _put_varint(len(block))

    index, cur, cur_bytes = [], [], 0
    for kv in items:
        cur.append(kv)
        cur_bytes += len(kv[0]) + len(kv[1]) + 3
        if cur_bytes >= block_size:
            index.append((cur[-1][0], emit(_build_block(cur))))         # separator = last key of the block
            cur, cur_bytes = [], 0
    if cur or not index:
        index.append((cur[-1][0] if cur else b"", emit(_build_block(cur))))
    meta = emit(_build_block([]))
    idx = emit(_build_block(index, restart_interval=1))
    foot = meta + idx
    out.extend(foot + b"\x00" * (40 - len(foot)) + struct.pack("<Q", TABLE_MAGIC))
    with open(path, "wb") as f:
        f.write(bytes(out))


# ---- tensor bundle -----------------------------------------------------------------------------------------
def _data_path(prefix, shard, nshards):
    return f"{prefix}.data-{shard:05d}-of-{nshards:05d}"


def read_tensor_bundle(prefix: str, verify_table: bool = True, verify_tensors: bool = False) -> Dict[str, np.ndarray]:
    """All numeric tensors of the bundle <prefix>.index / <prefix>.data-*; string tensors are skipped."""
    entries = read_table(prefix + ".index", verify_table)
    nshards = 1
    out, shards = {}, {}
    for key, val in entries:
        if key == b"":
            for f, _, v in _pb_fields(val):
                if f == 1:
                    nshards = v
                elif f == 2 and v != 0:
                    raise ValueError("big-endian tensor bundles are not supported")
            continue
        e = _parse_entry(val)
        if e["dtype"] == DT_STRING or e["sliced"]:
            continue
        if e["dtype"] not in DTYPES:
            raise ValueError(f"{key!r}: tensor dtype enum {e['dtype']} not supported")
        sid = e["shard_id"]
        if sid not in shards:
            shards[sid] = np.memmap(_data_path(prefix, sid, nshards), dtype=np.uint8, mode="r")
        raw = shards[sid][e["offset"]: e["offset"] + e["size"]]
        dt = DTYPES[e["dtype"]]
        count = int(np.prod(e["shape"])) if e["shape"] else 1
        if count * dt.itemsize != e["size"]:
            raise ValueError(f"{key!r}: {e['size']} bytes for shape {e['shape']} of {dt}")
        if verify_tensors and e["crc32c"] is not None and mask_crc(crc32c(raw.tobytes())) != e["crc32c"]:
            raise ValueError(f"{key!r}: tensor crc mismatch")
        out[key.decode()] = np.frombuffer(raw.tobytes(), dtype=dt).reshape(e["shape"]).copy()
    return out


def write_tensor_bundle(prefix: str, tensors: Dict[str, np.ndarray]):
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    header = b"\x08\x01" + b"\x1a\x02\x08\x01"          # num_shards = 1, little endian (default), version.producer = 1
    items = [(b"", header)]
    off = 0
    with open(_data_path(prefix, 0, 1), "wb") as f:
        for name in sorted(tensors):
            if isinstance(tensors[name], (bytes, bytearray)):                       # scalar DT_STRING tensor
                raw, crc = _string_tensor_bytes([bytes(tensors[name])])
                f.write(raw)
                items.append((name.encode(), _entry_proto(DT_STRING, (), off, len(raw), crc)))
                off += len(raw)
                continue
            a = np.asarray(tensors[name], order="C")       # (ascontiguousarray would turn scalars into [1])
            dt = a.dtype.newbyteorder("<") if a.dtype.byteorder == ">" else a.dtype
            a = a.astype(dt, copy=False)
            raw = a.tobytes()
            f.write(raw)
            items.append((name.encode(), _entry_proto(DTYPE_IDS[np.dtype(dt)], a.shape, off, len(raw),
                                                      mask_crc(crc32c(raw)))))
            off += len(raw)
    write_table(prefix + ".index", items)


# ---- string tensors and the object graph -----------------------------------------------------------------------
OBJECT_GRAPH_KEY = "_CHECKPOINTABLE_OBJECT_GRAPH"


def _string_tensor_bytes(values: List[bytes]) -> Tuple[bytes, int]:
    """DT_STRING layout of tensor_bundle.cc (WriteStringTensor): [varint64 length]* | masked crc32c of the lengths
    (each taken as a little-endian uint32) | the strings.  Returns (bytes, crc32c over lengths-as-uint32, the 4
    checksum bytes and the string bytes) — the value BundleEntryProto.crc32c stores, masked."""
    lengths = b"".join(_put_varint(len(v)) for v in values)
    crc = 0
    for v in values:
        crc = crc32c(struct.pack("<I", len(v)), crc)
    cks = struct.pack("<I", mask_crc(crc))
    crc = crc32c(cks, crc)
    for v in values:
        crc = crc32c(v, crc)
    return lengths + cks + b"".join(values), mask_crc(crc)


def _parse_string_tensor(raw: bytes, count: int) -> List[bytes]:
    pos, lens = 0, []
    for _ in range(count):
        n, pos = _get_varint(raw, pos)
        lens.append(n)
    crc = 0
    for n in lens:
        crc = crc32c(struct.pack("<I", n), crc)
    if struct.unpack("<I", raw[pos: pos + 4])[0] != mask_crc(crc):
        raise ValueError("string tensor: checksum of the lengths does not match")
    pos += 4
    out = []
    for n in lens:
        out.append(bytes(raw[pos: pos + n]))
        pos += n
    return out


def _pb_str(field: int, v: bytes) -> bytes:
    return _put_varint(field << 3 | 2) + _put_varint(len(v)) + v


def _pb_int(field: int, v: int) -> bytes:
    return _put_varint(field << 3) + _put_varint(v) if v else b""        # proto3: zero is the default, not written


def build_object_graph(keys: Iterable[str]) -> bytes:
    """Serialized TrackableObjectGraph (tensorflow/core/protobuf/trackable_object_graph.proto) for the variables whose
    checkpoint keys are `keys` (with the /.ATTRIBUTES/VARIABLE_VALUE suffix).  Node 0 is the root
    (`tf.train.Checkpoint(transformer=..., optimizer=...)`, core/models.py:322-323); every path component of a key is
    a child edge named like the Python attribute / list index; the object a key ends at carries
    SerializedTensor{name: 'VARIABLE_VALUE', full_name, checkpoint_key}.  Adam's moments
    (`<variable path>/.OPTIMIZER_SLOT/optimizer/<m|v>`) are nodes of their own, reachable only through the optimizer
    node's slot_variables table, as TensorFlow writes them."""
    nodes = [{"children": {}, "attrs": [], "slots": []}]

    def walk(path):
        n = 0
        for part in path:
            kids = nodes[n]["children"]
            if part not in kids:
                nodes.append({"children": {}, "attrs": [], "slots": []})
                kids[part] = len(nodes) - 1
            n = kids[part]
        return n

    slots = []
    for key in sorted(keys):
        assert key.endswith(SUFFIX), key
        path = key[: -len(SUFFIX)]
        if "/.OPTIMIZER_SLOT/" in path:
            slots.append((path, key))
            continue
        n = walk(path.split("/"))
        nodes[n]["attrs"].append(("VARIABLE_VALUE", path.split("/", 1)[-1], key))
    for path, key in slots:
        var_path, rest = path.split("/.OPTIMIZER_SLOT/")
        opt_name, slot_name = rest.split("/")
        var_node, opt_node = walk(var_path.split("/")), walk([opt_name])
        nodes.append({"children": {}, "attrs": [("VARIABLE_VALUE", f"{var_path.split('/', 1)[-1]}/{slot_name}", key)],
                      "slots": []})
        nodes[opt_node]["slots"].append((var_node, slot_name, len(nodes) - 1))
    out = b""
    for nd in nodes:
        body = b""
        for name, nid in nd["children"].items():
            body += _pb_str(1, _pb_int(1, nid) + _pb_str(2, name.encode()))
        for name, full, key in nd["attrs"]:
            body += _pb_str(2, _pb_str(1, name.encode()) + _pb_str(2, full.encode()) + _pb_str(3, key.encode()))
        for orig, slot_name, snode in nd["slots"]:
            body += _pb_str(3, _pb_int(1, orig) + _pb_str(2, slot_name.encode()) + _pb_int(3, snode))
        out += _pb_str(1, body)
    return out


def parse_object_graph(buf: bytes) -> List[dict]:
    """-> [{'children': {local_name: node_id}, 'attributes': [(name, full_name, checkpoint_key)],
            'slot_variables': [(original_variable_node_id, slot_name, slot_variable_node_id)]}] in node order."""
    nodes = []
    for f, _, v in _pb_fields(buf):
        if f != 1:
            continue
        nd = {"children": {}, "attributes": [], "slot_variables": []}
        for f2, _, v2 in _pb_fields(v):
            fields = {f3: v3 for f3, _, v3 in _pb_fields(v2)}
            if f2 == 1:
                nd["children"][bytes(fields.get(2, b"")).decode()] = int(fields.get(1, 0))
            elif f2 == 2:
                nd["attributes"].append(tuple(bytes(fields.get(i, b"")).decode() for i in (1, 2, 3)))
            elif f2 == 3:
                nd["slot_variables"].append((int(fields.get(1, 0)), bytes(fields.get(2, b"")).decode(), int(fields.get(3, 0))))
        nodes.append(nd)
    return nodes


def read_object_graph(prefix: str) -> List[dict]:
    """The parsed `_CHECKPOINTABLE_OBJECT_GRAPH` entry of a bundle, or [] when it has none."""
    for key, val in read_table(prefix + ".index"):
        if key.decode() == OBJECT_GRAPH_KEY:
            e = _parse_entry(val)
            if e["dtype"] != DT_STRING:
                raise ValueError("object graph entry is not a string tensor")
            with open(_data_path(prefix, e["shard_id"], _num_shards(prefix)), "rb") as f:
                f.seek(e["offset"])
                raw = f.read(e["size"])
            count = int(np.prod(e["shape"])) if e["shape"] else 1
            return parse_object_graph(_parse_string_tensor(raw, count)[0])
    return []


def _num_shards(prefix: str) -> int:
    for key, val in read_table(prefix + ".index"):
        if key == b"":
            for f, _, v in _pb_fields(val):
                if f == 1:
                    return int(v)
    return 1


def object_graph_variables(nodes: List[dict]) -> Dict[str, str]:
    """{attribute path from the root ('transformer/encoder/embedding/embeddings'): checkpoint_key} — what
    `Checkpoint.restore` finds by walking the children; slot variables appear as
    '<variable path>/.OPTIMIZER_SLOT/<optimizer child name>/<slot>'."""
    out, paths = {}, {0: ""}
    todo = [0]
    while todo:
        n = todo.pop()
        for name, nid in nodes[n]["children"].items():
            if nid not in paths:
                paths[nid] = (paths[n] + "/" + name).lstrip("/")
                todo.append(nid)
    for n, path in paths.items():
        for name, _, key in nodes[n]["attributes"]:
            out[path] = key
        for orig, slot_name, snode in nodes[n]["slot_variables"]:
            for _, _, key in nodes[snode]["attributes"]:
                out[f"{paths[orig]}/.OPTIMIZER_SLOT/{path}/{slot_name}"] = key
    return out


def latest_checkpoint(directory: str):
    """tf.train.latest_checkpoint: the `checkpoint` state file names the newest prefix (text proto)."""
    state = os.path.join(directory, "checkpoint")
    if not os.path.exists(state):
        return None
    with open(state) as f:
        for line in f:
            if line.startswith("model_checkpoint_path:"):
                name = line.split(":", 1)[1].strip().strip('"')
                return name if os.path.isabs(name) else os.path.join(directory, name)
    return None


# ---- object-graph keys of the reference model <-> parameter names of init.py ------------------------------------
def key_map(hps: dict) -> Dict[str, str]:
    """{parameter name (init.py): checkpoint key without the /.ATTRIBUTES/VARIABLE_VALUE suffix}.
    Attribute paths follow models/sketchformer.py:70-104 and builders/layers/transformer.py:154-158, 206-210,
    232-239, 276-283, 315-321, 363; `ffn` is a keras Sequential (layer_with_weights-0/1), the SelfAttn weights
    hang off the attributes W / b / V, and V2's extra Dense is spelled `embeding_layer` (transformer.py:100)."""
    L = hps["num_layers"]
    u = hps["lowerdim"]
    continuous = bool(hps.get("use_continuous_data", False))
    v2 = hps.get("attn_version", 1) != 1
    do_class = hps.get("do_classification", True) and bool(u)
    do_recon = hps.get("do_reconstruction", True)
    m: Dict[str, str] = {}
    root = "transformer/"

    def dense(ours, theirs):
        m[ours + ".k"] = root + theirs + "/kernel"
        m[ours + ".b"] = root + theirs + "/bias"

    def mha(ours, theirs):
        for a, b in (("wq", "wq"), ("wk", "wk"), ("wv", "wv"), ("wo", "dense")):
            dense(f"{ours}.{a}", f"{theirs}/{b}")

    def ln(ours, theirs):
        m[ours + ".g"] = root + theirs + "/gamma"
        m[ours + ".b"] = root + theirs + "/beta"

    def embed(ours, theirs):
        if continuous:
            dense(ours, theirs)
        else:
            m[ours] = root + theirs + "/embeddings"

    embed("enc.emb", "encoder/embedding")
    for i in range(L):
        t = f"encoder/enc_layers/{i}"
        mha(f"enc.l{i}.mha", t + "/mha")
        dense(f"enc.l{i}.ffn.w1", t + "/ffn/layer_with_weights-0")
        dense(f"enc.l{i}.ffn.w2", t + "/ffn/layer_with_weights-1")
        ln(f"enc.l{i}.ln1", t + "/layernorm1")
        ln(f"enc.l{i}.ln2", t + "/layernorm2")
    if u:
        for w in ("W", "b", "V"):
            m["pool." + w] = root + "bottleneck_layer/" + w
        if v2:
            dense("pool.emb", "bottleneck_layer/embeding_layer")
    if do_class:
        for j in range(int(hps.get("class_buffer_layers", 0))):
            dense(f"cbuf{j}", f"class_buffer/{j}")
        dense("cls", "classify_layer")
    if do_recon:
        if u:
            dense("exp", "expand_layer/expand_layer")
        embed("dec.emb", "decoder/embedding")
        for i in range(L):
            t = f"decoder/dec_layers/{i}"
            mha(f"dec.l{i}.mha1", t + "/mha1")
            mha(f"dec.l{i}.mha2", t + "/mha2")
            dense(f"dec.l{i}.ffn.w1", t + "/ffn/layer_with_weights-0")
            dense(f"dec.l{i}.ffn.w2", t + "/ffn/layer_with_weights-1")
            for j in (1, 2, 3):
                ln(f"dec.l{i}.ln{j}", f"{t}/layernorm{j}")
        dense("out", "output_layer")
    return m


STEP_KEYS = ("transformer/current_step", "optimizer/iter")


def import_checkpoint(prefix: str, hps: dict):
    """-> (params, adam_m, adam_v, info): dicts keyed by init.py names (float32 numpy, Keras shapes); the Adam
    dicts are empty when the bundle holds no optimizer slots; info = {'current_step', 'optimizer_iter'}."""
    bundle = read_tensor_bundle(prefix)
    params, adam_m, adam_v = {}, {}, {}
    missing = []
    for ours, theirs in key_map(hps).items():
        k = theirs + SUFFIX
        if k not in bundle:
            missing.append(k)
            continue
        params[ours] = bundle[k].astype(np.float32)
        for slot, dst in (("m", adam_m), ("v", adam_v)):
            ks = f"{theirs}/.OPTIMIZER_SLOT/optimizer/{slot}{SUFFIX}"
            if ks in bundle:
                dst[ours] = bundle[ks].astype(np.float32)
    if missing:
        raise KeyError(f"{prefix}: {len(missing)} variables of this model are not in the checkpoint, e.g. {missing[:3]}; "
                       f"it holds e.g. {sorted(k for k in bundle if k.startswith('transformer/'))[:3]}")
    info = {}
    for name, k in (("current_step", STEP_KEYS[0]), ("optimizer_iter", STEP_KEYS[1])):
        if k + SUFFIX in bundle:
            info[name] = int(bundle[k + SUFFIX].reshape(-1)[0])
    return params, adam_m, adam_v, info


def export_checkpoint(prefix: str, hps: dict, params: dict, adam_m: dict = None, adam_v: dict = None,
                      current_step: int = None, optimizer_iter: int = None):
    """Write parameters (init.py names, Keras shapes) under the reference's checkpoint keys."""
    km = key_map(hps)
    out = {}
    for ours, theirs in km.items():
        out[theirs + SUFFIX] = np.asarray(params[ours], dtype=np.float32)
        for slot, src in (("m", adam_m), ("v", adam_v)):
            if src:
                out[f"{theirs}/.OPTIMIZER_SLOT/optimizer/{slot}{SUFFIX}"] = np.asarray(src[ours], dtype=np.float32)
    if current_step is not None:
        out[STEP_KEYS[0] + SUFFIX] = np.asarray(current_step, dtype=np.int32)
    if optimizer_iter is not None:
        out[STEP_KEYS[1] + SUFFIX] = np.asarray(optimizer_iter, dtype=np.int64)
    out[OBJECT_GRAPH_KEY] = build_object_graph(out.keys())
    write_tensor_bundle(prefix, out)
    with open(os.path.join(os.path.dirname(os.path.abspath(prefix)), "checkpoint"), "w") as f:
        base = os.path.basename(prefix)
        f.write(f'model_checkpoint_path: "{base}"\nall_model_checkpoint_paths: "{base}"\n')
