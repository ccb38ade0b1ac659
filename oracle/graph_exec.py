"""ORACLE PIN (test infrastructure): execute the reference's OWN serialized TensorFlow graphs node by node in numpy.

The reference cannot run here (no TensorFlow), ships no tests and no golden vectors.  What it does ship is the
computation itself, serialized: ``embedders/yamnet_k2/models/yamnet_{wholehop,halfhop}/saved_model.pb`` and
``models/model_general_v3/saved_model.pb``.  This module parses those protos (tools/tfbundle.py) and interprets the
``__inference__wrapped_model_*`` function op by op -- every Pad / frame-gather / RFFT / MatMul / Conv2D /
FusedBatchNormV3 node, with the attributes and captured constants stored in the file -- so the restatement in
oracle/yamnet_oracle.py is checked against the reference's graph rather than against our reading of its Python.

It needs /root/reference, so it runs only in the build container: tools/make_golden.py uses it to write the
fixtures under tests/golden/ that travel to the GPU box.

Limits (stated, not hidden): the interpreter's kernels are numpy/scipy/torch-CPU, not TensorFlow's Eigen/oneDNN, so
float32 rounding differs at the 1e-6 level; YAMNet weights are whatever the caller passes (the blob is absent from
the reference checkout); the head weights are the reference's real ones.
"""
from __future__ import annotations

import os
import sys

import numpy as np

_TOOLS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools")
if _TOOLS not in sys.path:
    sys.path.insert(0, _TOOLS)
import tfbundle  # noqa: E402

_NP = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64, 10: np.bool_, 8: np.complex64}


def _tensor_to_np(t):
    shape = tuple(d.size for d in t.tensor_shape.dim)
    dt = _NP[t.dtype]
    if t.tensor_content:
        return np.frombuffer(t.tensor_content, dtype=dt).reshape(shape).copy()
    vals = {1: t.float_val, 2: t.double_val, 3: t.int_val, 9: t.int64_val, 10: t.bool_val}[t.dtype]
    a = np.array(list(vals), dtype=dt)
    n = int(np.prod(shape)) if shape else 1
    if a.size == 0:
        a = np.zeros(n, dtype=dt)
    elif a.size == 1 and n > 1:
        a = np.full(n, a[0], dtype=dt)
    return a.reshape(shape)


def _same_pad(size, stride, k):
    out = -(-size // stride)
    total = max((out - 1) * stride + k - size, 0)
    return total // 2, total - total // 2


def _strided_slice(x, begin, end, strides, attr):
    bm = attr["begin_mask"].i
    em = attr["end_mask"].i
    sm = attr["shrink_axis_mask"].i
    assert attr["ellipsis_mask"].i == 0 and attr["new_axis_mask"].i == 0
    idx = []
    for d in range(len(begin)):
        if sm & (1 << d):
            idx.append(int(begin[d]))
            continue
        b = None if bm & (1 << d) else int(begin[d])
        e = None if em & (1 << d) else int(end[d])
        idx.append(slice(b, e, int(strides[d])))
    return np.asarray(x[tuple(idx)])


class GraphFunction:
    """One library function of a SavedModel MetaGraph, runnable on numpy inputs."""

    def __init__(self, saved_model_path: str, prefix: str = "__inference__wrapped_model"):
        self.mg = tfbundle.read_meta_graphs(saved_model_path)[0]
        fns = [f for f in self.mg.graph_def.library.function if f.signature.name.startswith(prefix)]
        assert len(fns) == 1, [f.signature.name for f in fns]
        self.fn = fns[0]
        self.arg_names = [a.name for a in self.fn.signature.input_arg]
        # the call site in the top-level graph tells which constant / variable feeds which argument
        call = None
        for n in self.mg.graph_def.node:
            if n.op == "StatefulPartitionedCall" and len(n.input) == len(self.arg_names):
                call = n
                break
        assert call is not None
        self.call_inputs = [i.split(":")[0] for i in call.input]
        self.top = {n.name: n for n in self.mg.graph_def.node}
        self.ops_run = {}

    def bind(self, audio_or_input: np.ndarray, variables_in_order: list[np.ndarray]) -> dict:
        """argument name -> value: placeholder, captured Const nodes, then variables (VarHandleOp) in call order."""
        env = {}
        vi = 0
        for arg, src in zip(self.arg_names, self.call_inputs):
            node = self.top[src]
            if node.op == "Placeholder":
                env[arg] = audio_or_input
            elif node.op == "Const":
                env[arg] = _tensor_to_np(node.attr["value"].tensor)
            elif node.op == "VarHandleOp":
                v = variables_in_order[vi]
                shape = tuple(d.size for d in node.attr["shape"].shape.dim)
                assert tuple(v.shape) == shape, (src, v.shape, shape)
                env[arg] = v
                vi += 1
            else:
                raise NotImplementedError(node.op)
        assert vi == len(variables_in_order), (vi, len(variables_in_order))
        return env

    def variable_names(self) -> list[str]:
        return [s for s in self.call_inputs if self.top[s].op == "VarHandleOp"]

    # ------------------------------------------------------------------ interpreter
    def run(self, env: dict) -> np.ndarray:
        import scipy.fft
        import torch
        import torch.nn.functional as F
        vals = dict(env)

        def get(ref: str):
            if ref.startswith("^"):
                return None
            parts = ref.split(":")
            name = parts[0]
            if name in vals and len(parts) == 1:
                return vals[name]
            idx = int(parts[2]) if len(parts) == 3 else 0
            v = vals[name]
            return v[idx] if isinstance(v, tuple) else v

        for n in self.fn.node_def:
            op = n.op
            self.ops_run[op] = self.ops_run.get(op, 0) + 1
            a = n.attr
            i = [get(r) for r in n.input if not r.startswith("^")]
            if op == "Const":
                out = _tensor_to_np(a["value"].tensor)
            elif op in ("Identity", "ReadVariableOp"):
                out = i[0]
            elif op == "NoOp":
                continue
            elif op == "Shape":
                out = np.array(np.shape(i[0]), dtype=np.int32)
            elif op == "StridedSlice":
                out = _strided_slice(i[0], i[1], i[2], i[3], a)
            elif op == "Maximum":
                out = np.maximum(i[0], i[1])
            elif op == "Sub":
                out = np.subtract(i[0], i[1])
            elif op == "AddV2":
                out = np.add(i[0], i[1])
            elif op == "Mul":
                out = np.multiply(i[0], i[1])
            elif op == "RealDiv":
                out = np.divide(i[0], i[1]).astype(i[0].dtype)
            elif op == "FloorDiv":
                out = np.floor_divide(i[0], i[1])
            elif op == "FloorMod":
                out = np.mod(i[0], i[1])
            elif op == "Cast":
                out = np.asarray(i[0]).astype(_NP[a["DstT"].type])
            elif op == "Ceil":
                out = np.ceil(i[0])
            elif op == "Cos":
                out = np.cos(i[0], dtype=i[0].dtype)
            elif op == "Log":
                out = np.log(i[0])
            elif op == "Relu":
                out = np.maximum(i[0], 0)
            elif op == "Pack":
                out = np.stack([np.asarray(x) for x in i], axis=a["axis"].i)
            elif op == "ConcatV2":
                out = np.concatenate([np.atleast_1d(x) for x in i[:-1]], axis=int(i[-1]))
            elif op == "Range":
                out = np.arange(int(i[0]), int(i[1]), int(i[2]), dtype=np.asarray(i[0]).dtype)
            elif op == "Fill":
                out = np.full(tuple(int(x) for x in i[0]), i[1])
            elif op == "SplitV":
                sizes = [int(x) for x in i[1]]
                pts = np.cumsum(sizes)[:-1]
                out = tuple(np.split(i[0], pts, axis=int(i[2])))
            elif op == "Reshape":
                out = np.reshape(i[0], tuple(int(x) for x in np.atleast_1d(i[1])))
            elif op == "Pad":
                out = np.pad(i[0], [(int(p[0]), int(p[1])) for p in np.asarray(i[1]).reshape(-1, 2)])
            elif op == "GatherV2":
                out = np.take(i[0], i[1], axis=int(i[2]))
            elif op == "RFFT":
                nfft = int(np.atleast_1d(i[1])[0])
                out = scipy.fft.rfft(np.ascontiguousarray(i[0], dtype=np.float32), n=nfft, axis=-1).astype(np.complex64)
            elif op == "ComplexAbs":
                out = np.abs(i[0]).astype(np.float32)
            elif op == "MatMul":
                x, y = i[0], i[1]
                if a["transpose_a"].b:
                    x = x.T
                if a["transpose_b"].b:
                    y = y.T
                out = (x @ y).astype(np.float32)
            elif op == "BiasAdd":
                out = i[0] + i[1]
            elif op == "Mean":
                axes = tuple(int(x) for x in np.atleast_1d(i[1]))
                out = np.mean(i[0], axis=axes, keepdims=bool(a["keep_dims"].b), dtype=np.float32)
            elif op in ("Conv2D", "DepthwiseConv2dNative"):
                assert a["data_format"].s in (b"", b"NHWC")
                strides = list(a["strides"].list.i)
                pad = a["padding"].s.decode()
                x = torch.from_numpy(np.ascontiguousarray(i[0])).permute(0, 3, 1, 2)
                k = np.asarray(i[1])
                kh, kw = k.shape[0], k.shape[1]
                if pad == "SAME":
                    ph = _same_pad(x.shape[2], strides[1], kh)
                    pw = _same_pad(x.shape[3], strides[2], kw)
                    x = F.pad(x, (pw[0], pw[1], ph[0], ph[1]))
                else:
                    assert pad == "VALID"
                if op == "Conv2D":
                    w = torch.from_numpy(np.ascontiguousarray(k)).permute(3, 2, 0, 1)
                    y = F.conv2d(x, w, stride=(strides[1], strides[2]))
                else:
                    c = k.shape[2]
                    assert k.shape[3] == 1
                    w = torch.from_numpy(np.ascontiguousarray(k)).permute(2, 3, 0, 1)
                    y = F.conv2d(x, w, stride=(strides[1], strides[2]), groups=c)
                out = y.permute(0, 2, 3, 1).contiguous().numpy()
            elif op == "FusedBatchNormV3":
                assert not a["is_training"].b
                # SavedModels strip attributes that equal the op-def default; FusedBatchNormV3's default epsilon is
                # 1e-4, which is also what the reference's Params.batchnorm_epsilon asks for (params.py:48).
                eps = np.float32(a["epsilon"].f if "epsilon" in a else 1e-4)
                x, scale, offset, mean, var = i
                inv = (scale / np.sqrt(var + eps)).astype(np.float32)
                out = ((x - mean) * inv + offset).astype(np.float32)
                self.last_bn_epsilon = float(eps)
            else:
                raise NotImplementedError(f"op {op} ({n.name})")
            vals[n.name] = out
        ret = list(self.fn.ret.values())[0]
        return get(ret)


def variables_in_checkpoint_order(variables: dict) -> list[np.ndarray]:
    """YAMNet variables in the order the serving function captures them: kernel, beta, moving_mean,
    moving_variance per conv stage (layer_with_weights-0 .. -53)."""
    out = []
    idx = 0
    while True:
        base = f"layer_with_weights-{idx}"
        k = [n for n in (base + "/kernel", base + "/depthwise_kernel") if n in variables]
        if not k:
            break
        out.append(variables[k[0]])
        bn = f"layer_with_weights-{idx + 1}"
        out += [variables[bn + "/beta"], variables[bn + "/moving_mean"], variables[bn + "/moving_variance"]]
        idx += 2
    return out


def run_yamnet_graph(reference_root: str, samples: np.ndarray, variables: dict, hop: str = "wholehop"):
    g = GraphFunction(os.path.join(reference_root, f"embedders/yamnet_k2/models/yamnet_{hop}/saved_model.pb"))
    env = g.bind(np.asarray(samples, dtype=np.float32), variables_in_checkpoint_order(variables))
    return g.run(env), g


def run_head_graph(reference_root: str, embeddings: np.ndarray):
    """models/model_general_v3: the reference's real Dense(13) weights from its variables.data."""
    mdir = os.path.join(reference_root, "models/model_general_v3")
    g = GraphFunction(os.path.join(mdir, "saved_model.pb"))
    idx = {e.name: e for e in tfbundle.read_bundle_index(os.path.join(mdir, "variables/variables.index"))}
    data = open(os.path.join(mdir, "variables/variables.data-00000-of-00001"), "rb").read()

    def var(name):
        e = idx[name + "/.ATTRIBUTES/VARIABLE_VALUE"]
        return np.frombuffer(data[e.offset:e.offset + e.size], dtype="<f4").reshape(e.shape).copy()
    vs = [var("layer_with_weights-0/kernel"), var("layer_with_weights-0/bias")]
    env = g.bind(np.asarray(embeddings, dtype=np.float32), vs)
    return g.run(env), g
