"""TensorFlow tensor-bundle reader / writer (sketchformer_b200/tf_checkpoint.py): table format known answers,
round trips, and the reference's object-graph key layout (core/models.py:321-331)."""
import os
import struct

import numpy as np
import pytest

from sketchformer_b200 import tf_checkpoint as T
from sketchformer_b200.init import init_params

HPS = dict(num_layers=2, d_model=32, dff=64, num_heads=4, lowerdim=16, max_seq_len=12, vocab_size=50, n_classes=7)


def test_crc32c_known_answers():
    # RFC 3720 B.4 test vectors
    assert T.crc32c(b"123456789") == 0xE3069283
    assert T.crc32c(b"\x00" * 32) == 0x8A9136AA
    assert T.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert T.crc32c(bytes(range(32))) == 0x46DD794E
    # LevelDB mask: rotate right 15, add constant
    assert T.mask_crc(0) == 0xA282EAD8
    assert T.mask_crc(0xE3069283) == ((((0xE3069283 >> 15) | (0xE3069283 << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def test_snappy_known_answers():
    # literal "abcd", then copy (1-byte offset form) of 4 bytes from offset 4, then a self-overlapping run
    assert T.snappy_uncompress(bytes([8, 0x0C]) + b"abcd" + bytes([0x01, 0x04])) == b"abcdabcd"
    assert T.snappy_uncompress(bytes([9, 0x00]) + b"x" + bytes([(7 << 2) | 2, 1, 0])) == b"x" * 9
    with pytest.raises(ValueError):
        T.snappy_uncompress(bytes([4, 0x01, 0x09]))


def test_hand_assembled_table(tmp_path):
    """A table built byte by byte here (prefix compression, two data blocks), independent of write_table()."""
    def block(entries, restarts):
        b = b"".join(bytes([s, len(k), len(v)]) + k + v for s, k, v in entries)
        return b + b"".join(struct.pack("<I", r) for r in restarts) + struct.pack("<I", len(restarts))

    def framed(b):
        return b + b"\x00" + struct.pack("<I", T.mask_crc(T.crc32c(b + b"\x00")))

    b0 = block([(0, b"apple", b"1"), (3, b"ly", b"22"), (0, b"bar", b"")], [0, 15])   # apple, apply, bar
    b1 = block([(0, b"cat", b"333")], [0])
    meta = block([], [0])
    out = framed(b0)
    o1 = len(out); out += framed(b1)
    om = len(out); out += framed(meta)
    idx = block([(0, b"bar", bytes([0, len(b0)])), (0, b"d", bytes([o1, len(b1)]))], [0, 6])
    oi = len(out); out += framed(idx)
    foot = bytes([om, len(meta), oi, len(idx)])
    out += foot + b"\x00" * (40 - len(foot)) + struct.pack("<Q", T.TABLE_MAGIC)
    p = tmp_path / "t.index"
    p.write_bytes(out)
    assert T.read_table(str(p)) == [(b"apple", b"1"), (b"apply", b"22"), (b"bar", b""), (b"cat", b"333")]
    bad = bytearray(out); bad[2] ^= 1
    p.write_bytes(bytes(bad))
    with pytest.raises(ValueError):
        T.read_table(str(p))


def test_bundle_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    tensors = {f"transformer/layer/{i}/kernel{T.SUFFIX}": rng.standard_normal((i + 1, 3)).astype(np.float32)
               for i in range(300)}                                   # several index blocks
    tensors["optimizer/iter" + T.SUFFIX] = np.asarray(1234567890123, dtype=np.int64)
    tensors["scalar_f64"] = np.asarray(2.5)
    tensors["empty"] = np.zeros((0, 4), dtype=np.float32)
    prefix = str(tmp_path / "ckpt-3")
    T.write_tensor_bundle(prefix, tensors)
    back = T.read_tensor_bundle(prefix, verify_tensors=True)
    assert sorted(back) == sorted(tensors)
    for k, v in tensors.items():
        assert back[k].dtype == v.dtype and back[k].shape == v.shape and np.array_equal(back[k], v), k
    # the entry of one tensor, decoded by hand: dtype DT_FLOAT, dims, size
    entries = dict(T.read_table(prefix + ".index"))
    e = T._parse_entry(entries[f"transformer/layer/4/kernel{T.SUFFIX}".encode()])
    assert e["dtype"] == 1 and e["shape"] == [5, 3] and e["size"] == 60 and e["shard_id"] == 0
    # a flipped data byte is caught by the tensor crc
    with open(prefix + ".data-00000-of-00001", "r+b") as f:
        f.seek(5); b = f.read(1); f.seek(5); f.write(bytes([b[0] ^ 0x40]))
    with pytest.raises(ValueError):
        T.read_tensor_bundle(prefix, verify_tensors=True)


@pytest.mark.parametrize("extra", [dict(), dict(use_continuous_data=True), dict(attn_version=2, class_buffer_layers=2),
                                   dict(lowerdim=0, do_classification=False), dict(do_reconstruction=False)])
def test_key_map_covers_every_variable(extra):
    hps = dict(HPS, **extra)
    params = init_params(hps, 0)
    km = T.key_map(hps)
    assert sorted(km) == sorted(params)
    assert len(set(km.values())) == len(km)
    if not extra:
        assert km["enc.l1.mha.wo.k"] == "transformer/encoder/enc_layers/1/mha/dense/kernel"
        assert km["dec.l0.ffn.w2.b"] == "transformer/decoder/dec_layers/0/ffn/layer_with_weights-1/bias"
        assert km["dec.l1.ln3.g"] == "transformer/decoder/dec_layers/1/layernorm3/gamma"
        assert km["pool.V"] == "transformer/bottleneck_layer/V"
        assert km["exp.k"] == "transformer/expand_layer/expand_layer/kernel"
        assert km["enc.emb"] == "transformer/encoder/embedding/embeddings"


def test_checkpoint_export_import(tmp_path):
    params = {k: v.numpy() for k, v in init_params(HPS, 3).items()}
    m = {k: v * 0.5 for k, v in params.items()}
    v = {k: v * v for k, v in params.items()}
    prefix = str(tmp_path / "weights" / "ckpt-7")
    T.export_checkpoint(prefix, HPS, params, m, v, current_step=699, optimizer_iter=700)
    assert T.latest_checkpoint(str(tmp_path / "weights")) == prefix
    p2, m2, v2, info = T.import_checkpoint(prefix, HPS)
    assert info == {"current_step": 699, "optimizer_iter": 700}
    for k in params:
        assert np.array_equal(p2[k], params[k]) and np.array_equal(m2[k], m[k]) and np.array_equal(v2[k], v[k])
    with pytest.raises(KeyError):
        T.import_checkpoint(prefix, dict(HPS, num_layers=3))


def test_string_tensor_layout():
    """tensor_bundle.cc WriteStringTensor: varint lengths | masked crc32c of the lengths as uint32 | bytes."""
    vals = [b"", b"abc", b"x" * 300]
    raw, crc = T._string_tensor_bytes(vals)
    assert raw[:4] == b"\x00\x03\xac\x02"                                     # 0, 3, 300 as varints
    want = 0
    for n in (0, 3, 300):
        want = T.crc32c(struct.pack("<I", n), want)
    assert struct.unpack("<I", raw[4:8])[0] == T.mask_crc(want)
    assert raw[8:] == b"abc" + b"x" * 300
    assert T._parse_string_tensor(raw, 3) == vals
    assert crc == T.mask_crc(T.crc32c(raw[8:], T.crc32c(raw[4:8], want)))
    with pytest.raises(ValueError):
        T._parse_string_tensor(raw[:4] + b"\0\0\0\0" + raw[8:], 3)


def _graph_message_class():
    """TrackableObjectGraph declared to the protobuf runtime (schema restated from
    tensorflow/core/protobuf/trackable_object_graph.proto): an encoder independent of tf_checkpoint's hand-rolled one."""
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    F = descriptor_pb2.FieldDescriptorProto
    fd = descriptor_pb2.FileDescriptorProto(name="tog_test.proto", package="togtest", syntax="proto3")
    g = fd.message_type.add(name="TrackableObjectGraph")
    o = g.nested_type.add(name="TrackableObject")
    r = o.nested_type.add(name="ObjectReference")
    r.field.add(name="node_id", number=1, type=F.TYPE_INT32, label=F.LABEL_OPTIONAL)
    r.field.add(name="local_name", number=2, type=F.TYPE_STRING, label=F.LABEL_OPTIONAL)
    t = o.nested_type.add(name="SerializedTensor")
    for i, n in enumerate(("name", "full_name", "checkpoint_key"), 1):
        t.field.add(name=n, number=i, type=F.TYPE_STRING, label=F.LABEL_OPTIONAL)
    v = o.nested_type.add(name="SlotVariableReference")
    v.field.add(name="original_variable_node_id", number=1, type=F.TYPE_INT32, label=F.LABEL_OPTIONAL)
    v.field.add(name="slot_name", number=2, type=F.TYPE_STRING, label=F.LABEL_OPTIONAL)
    v.field.add(name="slot_variable_node_id", number=3, type=F.TYPE_INT32, label=F.LABEL_OPTIONAL)
    pre = ".togtest.TrackableObjectGraph.TrackableObject."
    o.field.add(name="children", number=1, type=F.TYPE_MESSAGE, label=F.LABEL_REPEATED, type_name=pre + "ObjectReference")
    o.field.add(name="attributes", number=2, type=F.TYPE_MESSAGE, label=F.LABEL_REPEATED, type_name=pre + "SerializedTensor")
    o.field.add(name="slot_variables", number=3, type=F.TYPE_MESSAGE, label=F.LABEL_REPEATED,
                type_name=pre + "SlotVariableReference")
    g.field.add(name="nodes", number=1, type=F.TYPE_MESSAGE, label=F.LABEL_REPEATED,
                type_name=".togtest.TrackableObjectGraph.TrackableObject")
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName("togtest.TrackableObjectGraph"))


def test_object_graph_entry(tmp_path):
    """export_checkpoint writes `_CHECKPOINTABLE_OBJECT_GRAPH`: walking its children from the root
    (`transformer` / `optimizer`, core/models.py:322-323) reaches every variable of the bundle under its own key, the
    Adam moments hang off the optimizer node's slot table, and the protobuf runtime reads the same message."""
    params = {k: v.numpy() for k, v in init_params(HPS, 3).items()}
    m = {k: v * 0.5 for k, v in params.items()}
    v = {k: v * v for k, v in params.items()}
    prefix = str(tmp_path / "w" / "ckpt-3")
    T.export_checkpoint(prefix, HPS, params, m, v, current_step=12, optimizer_iter=13)
    bundle = T.read_tensor_bundle(prefix, verify_tensors=True)            # the string entry does not disturb the numeric reader
    assert T.OBJECT_GRAPH_KEY not in bundle
    nodes = T.read_object_graph(prefix)
    assert set(nodes[0]["children"]) == {"transformer", "optimizer"}
    found = T.object_graph_variables(nodes)
    assert {k + T.SUFFIX for k in found} == set(bundle) and all(found[k] == k + T.SUFFIX for k in found)
    tr = nodes[nodes[0]["children"]["transformer"]]
    enc = nodes[tr["children"]["encoder"]]
    assert sorted(nodes[enc["children"]["enc_layers"]]["children"]) == [str(i) for i in range(HPS["num_layers"])]
    opt = nodes[nodes[0]["children"]["optimizer"]]
    assert set(opt["children"]) == {"iter"} and len(opt["slot_variables"]) == 2 * len(params)
    assert {s for _, s, _ in opt["slot_variables"]} == {"m", "v"}
    # the same bytes through google.protobuf
    raw = None
    for key, val in T.read_table(prefix + ".index"):
        if key.decode() == T.OBJECT_GRAPH_KEY:
            e = T._parse_entry(val)
            assert e["dtype"] == T.DT_STRING and e["shape"] == []
            data = open(T._data_path(prefix, 0, 1), "rb").read()[e["offset"]: e["offset"] + e["size"]]
            raw = T._parse_string_tensor(data, 1)[0]
            assert T.mask_crc(T.crc32c(data[-len(raw):], T.crc32c(data[:len(data) - len(raw)][-4:],
                              T.crc32c(struct.pack("<I", len(raw)))))) == e["crc32c"]
    msg = _graph_message_class()()
    msg.ParseFromString(raw)
    assert len(msg.nodes) == len(nodes)
    for a, b in zip(msg.nodes, nodes):
        assert {c.local_name: c.node_id for c in a.children} == b["children"]
        assert [(t.name, t.full_name, t.checkpoint_key) for t in a.attributes] == b["attributes"]
        assert [(x.original_variable_node_id, x.slot_name, x.slot_variable_node_id) for x in a.slot_variables] == b["slot_variables"]
    assert msg.SerializeToString() == raw             # canonical field order: byte-identical re-encoding
    # a bundle without the entry (e.g. one written by name only) parses to an empty graph
    T.write_tensor_bundle(str(tmp_path / "plain"), {"a": np.zeros(3, np.float32)})
    assert T.read_object_graph(str(tmp_path / "plain")) == []


# ---- checkpoint keys against the reference's attribute structure ---------------------------------------------------
def _tracked_variables(obj, prefix=""):
    """TF object-graph naming restated (tf.train.Checkpoint: a variable's key is the chain of attribute names that
    leads to it, list elements by index, the layers of a keras Sequential as `layer_with_weights-N`), applied to an
    object built by the REFERENCE's code on oracle/tf_shim.py: attribute names and nesting are the reference's."""
    import torch
    from oracle import tf_shim
    out = {}
    if isinstance(obj, tf_shim.Sequential):
        for n, lay in enumerate(obj.layers):
            out.update(_tracked_variables(lay, f"{prefix}/layer_with_weights-{n}"))
        return out
    for name, val in vars(obj).items():
        if name in ("pos_encoding",):                       # a constant tensor, not a variable
            continue
        path = f"{prefix}/{name}" if prefix else name
        if isinstance(val, torch.Tensor):
            out[path] = tuple(val.shape)
        elif isinstance(val, tf_shim.Layer):
            out.update(_tracked_variables(val, path))
        elif isinstance(val, list) and val and all(isinstance(v, tf_shim.Layer) for v in val):
            for i, v in enumerate(val):
                out.update(_tracked_variables(v, f"{path}/{i}"))
    return out


@pytest.mark.skipif(not os.path.exists("/root/reference/models/sketchformer.py"),
                    reason="/root/reference is not present on this box")
@pytest.mark.parametrize("variant", ["default", "v2_classbuffer", "lowerdim0", "continuous", "no_recon"])
def test_checkpoint_keys_follow_the_reference_attribute_structure(variant):
    """key_map() against the variables of a model that the reference's build_model() created (models/sketchformer.py
    :63-107 on the TF-API shim): same attribute paths under the `transformer/` root (core/models.py:322-323), same
    shapes as our parameters."""
    import torch
    from oracle import sketchformer_oracle as O
    from oracle import tf_shim
    from oracle.make_golden import TINY, TINY_LOWERDIM0, TINY_V2
    from oracle.reference_model import build_reference
    from sketchformer_b200 import synth
    kw = {"default": TINY, "v2_classbuffer": dict(TINY_V2), "lowerdim0": TINY_LOWERDIM0,
          "continuous": dict(TINY, use_continuous_data=True), "no_recon": dict(TINY, do_reconstruction=False)}[variant]
    cfg = O.OracleConfig(**kw)
    p = O.init_params(cfg, 0, torch.float64)
    if cfg.use_continuous_data:
        x = torch.from_numpy(synth.continuous_batch(2, cfg.max_seq_len, cfg.n_classes, 1)[0]).double()
    else:
        x = torch.from_numpy(synth.token_batch(2, cfg.max_seq_len, cfg.n_classes, 1, vocab=cfg.vocab_size)[0])
    m = build_reference(tf_shim.load_reference_model(), cfg, p, x)
    layers = {k: v for k, v in vars(m).items() if isinstance(v, (tf_shim.Layer, list))}
    holder = type("Holder", (), {})()
    holder.__dict__.update(layers)
    ref_vars = {"transformer/" + k: v for k, v in _tracked_variables(holder).items()}
    hps = dict(num_layers=cfg.num_layers, lowerdim=cfg.lowerdim, attn_version=cfg.attn_version,
               do_classification=cfg.do_classification, do_reconstruction=cfg.do_reconstruction,
               class_buffer_layers=cfg.class_buffer_layers, use_continuous_data=cfg.use_continuous_data)
    km = T.key_map(hps)
    assert set(km) == set(p), (set(km) ^ set(p))
    ours = {key: tuple(p[name].shape) for name, key in km.items()}
    assert set(ours) == set(ref_vars), sorted(set(ours) ^ set(ref_vars))
    for key in ours:
        assert ours[key] == ref_vars[key], key
