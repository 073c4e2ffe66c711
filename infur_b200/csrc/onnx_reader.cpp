// ONNX protobuf-wire reader + lowering to the fused conv list.  See onnx_reader.h.
#include "onnx_reader.h"

#include <climits>

#include <cstdio>
#include <cstring>
#include <set>
#include <sstream>
#include <stdexcept>

#include "../../include/infur_b200.h"

namespace infur {
namespace {

struct Cursor {
  const uint8_t* p;
  const uint8_t* end;
  bool done() const { return p >= end; }
  uint64_t varint() {
    uint64_t r = 0;
    int shift = 0;
    while (true) {
      if (p >= end) throw std::runtime_error("onnx: truncated varint");
      uint8_t c = *p++;
      r |= (uint64_t)(c & 0x7f) << shift;
      if (!(c & 0x80)) return r;
      shift += 7;
      if (shift > 63) throw std::runtime_error("onnx: varint too long");
    }
  }
  // Reads one field header; for length-delimited fields returns the sub-range.
  bool next(uint32_t& field, uint32_t& wire, uint64_t& val, Cursor& sub) {
    if (done()) return false;
    uint64_t key = varint();
    field = (uint32_t)(key >> 3);
    wire = (uint32_t)(key & 7);
    switch (wire) {
      case 0: val = varint(); break;
      case 1:
        if (end - p < 8) throw std::runtime_error("onnx: truncated fixed64");
        memcpy(&val, p, 8); p += 8; break;
      case 2: {
        uint64_t len = varint();
        if ((uint64_t)(end - p) < len) throw std::runtime_error("onnx: truncated length-delimited field");
        sub.p = p; sub.end = p + len; p += len; break;
      }
      case 5: {
        if (end - p < 4) throw std::runtime_error("onnx: truncated fixed32");
        uint32_t v; memcpy(&v, p, 4); val = v; p += 4; break;
      }
      default: throw std::runtime_error("onnx: unsupported wire type");
    }
    return true;
  }
};

std::string str(const Cursor& c) { return std::string((const char*)c.p, (size_t)(c.end - c.p)); }

void packed_varints(uint32_t wire, uint64_t val, Cursor sub, std::vector<int64_t>& out) {
  if (wire == 0) { out.push_back((int64_t)val); return; }
  while (!sub.done()) out.push_back((int64_t)sub.varint());
}

void parse_tensor(Cursor c, OnnxTensor& t) {
  uint32_t f, w; uint64_t v; Cursor s{};
  while (c.next(f, w, v, s)) {
    switch (f) {
      case 1: packed_varints(w, v, s, t.dims); break;
      case 2: t.dtype = (int32_t)v; break;
      case 4:
        if (w == 2) { size_t n = (size_t)(s.end - s.p) / 4; size_t o = t.float_data.size(); t.float_data.resize(o + n); memcpy(t.float_data.data() + o, s.p, n * 4); }
        else { float x; uint32_t u = (uint32_t)v; memcpy(&x, &u, 4); t.float_data.push_back(x); }
        break;
      case 5: { std::vector<int64_t> tmp; packed_varints(w, v, s, tmp); for (auto x : tmp) t.int32_data.push_back((int32_t)x); break; }
      case 7: packed_varints(w, v, s, t.int64_data); break;
      case 8: t.name = str(s); break;
      case 9: t.raw = s.p; t.raw_size = (size_t)(s.end - s.p); break;
      case 14: if (v != 0) throw std::runtime_error("onnx: external tensor data is not supported"); break;
      default: break;
    }
  }
}

void parse_attr(Cursor c, OnnxAttr& a) {
  uint32_t f, w; uint64_t v; Cursor s{};
  while (c.next(f, w, v, s)) {
    switch (f) {
      case 1: a.name = str(s); break;
      case 2: { uint32_t u = (uint32_t)v; memcpy(&a.f, &u, 4); break; }
      case 3: a.i = (int64_t)v; break;
      case 4: a.s = str(s); break;
      case 5: a.has_t = true; parse_tensor(s, a.t); break;
      case 8: packed_varints(w, v, s, a.ints); break;
      default: break;
    }
  }
}

void parse_node(Cursor c, OnnxNode& n) {
  uint32_t f, w; uint64_t v; Cursor s{};
  while (c.next(f, w, v, s)) {
    switch (f) {
      case 1: n.in.push_back(str(s)); break;
      case 2: n.out.push_back(str(s)); break;
      case 3: n.name = str(s); break;
      case 4: n.op = str(s); break;
      case 5: n.attrs.emplace_back(); parse_attr(s, n.attrs.back()); break;
      default: break;
    }
  }
}

void parse_value_info(Cursor c, OnnxValueInfo& vi) {
  uint32_t f, w; uint64_t v; Cursor s{};
  while (c.next(f, w, v, s)) {
    if (f == 1) vi.name = str(s);
    else if (f == 2 && w == 2) {           // TypeProto
      Cursor ty = s; Cursor s2{};
      while (ty.next(f, w, v, s2)) {
        if (f != 1 || w != 2) continue;    // tensor_type
        Cursor tt = s2; Cursor s3{};
        while (tt.next(f, w, v, s3)) {
          if (f == 1) vi.elem_type = (int32_t)v;
          else if (f == 2 && w == 2) {     // TensorShapeProto
            vi.has_shape = true;
            Cursor sh = s3; Cursor s4{};
            while (sh.next(f, w, v, s4)) {
              if (f != 1 || w != 2) continue;   // Dimension
              int64_t dim = -1;
              Cursor d = s4; Cursor s5{};
              while (d.next(f, w, v, s5)) if (f == 1 && w == 0) dim = (int64_t)v;
              vi.dims.push_back(dim);
            }
          }
        }
      }
    }
  }
}

void parse_graph(Cursor c, OnnxGraph& g) {
  uint32_t f, w; uint64_t v; Cursor s{};
  while (c.next(f, w, v, s)) {
    switch (f) {
      case 1: g.nodes.emplace_back(); parse_node(s, g.nodes.back()); break;
      case 5: { OnnxTensor t; parse_tensor(s, t); g.inits[t.name] = std::move(t); break; }
      case 11: g.inputs.emplace_back(); parse_value_info(s, g.inputs.back()); break;
      case 12: g.outputs.emplace_back(); parse_value_info(s, g.outputs.back()); break;
      default: break;
    }
  }
}

const char* dtype_name(int32_t t) {
  // Debug print of onnxruntime::TensorElementDataType, as shown by ModelInfo.input0_dtype
  switch (t) {
    case 1: return "Float"; case 2: return "Uint8"; case 3: return "Int8"; case 4: return "Uint16";
    case 5: return "Int16"; case 6: return "Int32"; case 7: return "Int64"; case 8: return "String";
    case 10: return "Float16"; case 11: return "Double"; case 12: return "Uint32"; case 13: return "Uint64";
    default: return "Undefined";
  }
}

// numel with every dimension checked: a malformed file must fail the load, not wrap around into a small count
size_t checked_numel(const OnnxTensor& t, const std::string& what) {
  size_t n = 1;
  for (int64_t d : t.dims) {
    if (d < 0 || d > (int64_t)INT32_MAX) throw ModelError(INFUR_E_MODEL_LOAD, what + ": initializer '" + t.name + "' has a negative or oversized dimension");
    if (d != 0 && n > (size_t)1 << 40) throw ModelError(INFUR_E_MODEL_LOAD, what + ": initializer '" + t.name + "' is too large");
    n *= (size_t)d;
  }
  return n;
}

// kernel size, stride and dilation positive and small, padding non-negative: build_plan divides by the stride
void check_geometry(const std::string& what, int64_t k, int64_t stride, int64_t dil, int64_t pad) {
  if (k < 1 || k > 64 || stride < 1 || stride > 64 || dil < 1 || dil > 64 || pad < 0 || pad > 4096)
    throw ModelError(INFUR_E_MODEL_LOAD, what + ": kernel size, stride and dilation must be 1..64 and padding 0..4096");
}

std::vector<float> tensor_f32(const OnnxTensor& t, const std::string& what) {
  if (t.dtype != 1) throw ModelError(INFUR_E_MODEL_LOAD, what + ": initializer '" + t.name + "' is not FLOAT (quantised models are not supported)");
  size_t n = checked_numel(t, what);
  std::vector<float> v(n);
  if (t.raw && t.raw_size == n * 4) memcpy(v.data(), t.raw, n * 4);
  else if (t.float_data.size() == n) v = t.float_data;
  else throw ModelError(INFUR_E_MODEL_LOAD, what + ": initializer '" + t.name + "' has inconsistent data size");
  return v;
}

// Integer initializer (UINT8 / INT8 / INT32; raw_data or int32_data) widened to int32.
std::vector<int32_t> tensor_i32(const OnnxTensor& t, const std::string& what) {
  const size_t n = checked_numel(t, what);
  std::vector<int32_t> v(n);
  const size_t esz = t.dtype == 6 ? 4 : (t.dtype == 2 || t.dtype == 3) ? 1 : 0;
  if (!esz) throw ModelError(INFUR_E_MODEL_LOAD, what + ": initializer '" + t.name + "' must be UINT8, INT8 or INT32");
  if (t.raw && t.raw_size == n * esz) {
    for (size_t i = 0; i < n; ++i) {
      if (t.dtype == 2) v[i] = t.raw[i];
      else if (t.dtype == 3) v[i] = (int8_t)t.raw[i];
      else memcpy(&v[i], t.raw + 4 * i, 4);
    }
  } else if (t.int32_data.size() == n) {
    v = t.int32_data;
  } else throw ModelError(INFUR_E_MODEL_LOAD, what + ": initializer '" + t.name + "' has inconsistent data size");
  return v;
}

int64_t attr_i(const OnnxNode& n, const char* name, int64_t dflt) {
  auto* a = n.attr(name);
  return a ? a->i : dflt;
}

std::vector<int64_t> attr_ints(const OnnxNode& n, const char* name, std::vector<int64_t> dflt) {
  auto* a = n.attr(name);
  return (a && !a->ints.empty()) ? a->ints : dflt;
}

bool all_eq(const std::vector<int64_t>& v, int64_t x) {
  for (auto e : v) if (e != x) return false;
  return true;
}

}  // namespace

void read_file(const std::string& path, std::vector<uint8_t>& bytes) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) throw std::runtime_error("cannot open '" + path + "'");
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  if (n < 0) { fclose(f); throw std::runtime_error("cannot stat '" + path + "'"); }
  bytes.resize((size_t)n);
  size_t got = n ? fread(bytes.data(), 1, (size_t)n, f) : 0;
  fclose(f);
  if (got != (size_t)n) throw std::runtime_error("short read on '" + path + "'");
}

void parse_onnx(std::vector<uint8_t>&& bytes, OnnxGraph& g) {
  g = OnnxGraph();
  g.file = std::move(bytes);
  Cursor c{g.file.data(), g.file.data() + g.file.size()};
  uint32_t f, w; uint64_t v; Cursor s{};
  bool has_graph = false;
  while (c.next(f, w, v, s)) {
    switch (f) {
      case 1: g.ir_version = (int64_t)v; break;
      case 2: if (w == 2) g.producer = str(s); break;
      case 7: if (w == 2) { parse_graph(s, g); has_graph = true; } break;
      case 8: {
        if (w != 2) break;
        Cursor o = s; Cursor s2{}; std::string domain; int64_t ver = 0;
        while (o.next(f, w, v, s2)) { if (f == 1) domain = str(s2); else if (f == 2) ver = (int64_t)v; }
        if (domain.empty() || domain == "ai.onnx") g.opset = ver;
        break;
      }
      default: break;
    }
  }
  if (!has_graph || g.nodes.empty()) throw std::runtime_error("onnx: no graph in file (not an ONNX model?)");
}

// ---------------------------------------------------------------------------------------------

void lower_model(const OnnxGraph& g, LoweredModel& m) {
  m = LoweredModel();
  // ---- inputs / outputs as ORT reports them: graph inputs minus initializers
  std::vector<const OnnxValueInfo*> real_inputs;
  for (auto& vi : g.inputs) if (!g.inits.count(vi.name)) real_inputs.push_back(&vi);
  if (real_inputs.empty()) throw ModelError(INFUR_E_MODEL_LOAD, "model has no inputs");
  for (auto* vi : real_inputs) m.io.input_names.push_back(vi->name);
  for (auto& vi : g.outputs) m.io.output_names.push_back(vi.name);
  const OnnxValueInfo& in0 = *real_inputs[0];
  m.io.input0_dtype = dtype_name(in0.elem_type);

  // infer_img_pre_proc (predict_onnx.rs:223-265), same checks in the same order
  int col_dim = -1;
  for (size_t i = 0; i < in0.dims.size(); ++i) if (in0.dims[i] == 3) { col_dim = (int)i; break; }
  if (col_dim < 0) throw ModelError(INFUR_E_MODEL_INPUT_FORMAT, "couldn't locate model's color input by dimension length 3");
  if (in0.dims.size() != 4)
    throw ModelError(INFUR_E_MODEL_INPUT_FORMAT, "only 4 dimensions supported got " + std::to_string(in0.dims.size()));
  if (col_dim == 1) m.io.nchw = true;
  else if (col_dim == 3) m.io.nchw = false;
  else throw ModelError(INFUR_E_MODEL_INPUT_FORMAT, "color dimension only at NCHW or NHWC but not in position " + std::to_string(col_dim) + " supported");
  if (in0.elem_type == 1) m.io.float_input = true;
  else if (in0.elem_type == 2) m.io.float_input = false;
  else throw ModelError(INFUR_E_MODEL_INPUT_FORMAT, std::string("only Float (f32) and Uint8 (u8) input supported, got ") + dtype_name(in0.elem_type));
  m.io.rgb = m.io.float_input;  // Model::control (:296-301): Float -> RGB + torchvision norm, else BGR

  // ---- pass 1: aliases, compute-node list, consumer counts
  static const std::set<std::string> shape_ops = {"Shape", "Gather", "Unsqueeze", "Concat", "Slice", "Cast", "Constant", "Squeeze", "Floor", "Mul", "Div"};
  std::map<std::string, std::string> alias;
  auto resolve = [&](const std::string& n) {
    std::string r = n;
    for (int guard = 0; guard < 64; ++guard) { auto it = alias.find(r); if (it == alias.end()) break; r = it->second; }
    return r;
  };
  std::vector<const OnnxNode*> compute;
  std::set<std::string> shape_values;  // outputs of the size-computing subgraph
  // Input adapters of NHWC / Uint8 models (infer_img_pre_proc's other conventions, predict_onnx.rs:240-262): ONNX Conv is
  // NCHW / float only, so such a model starts with Transpose(perm 0,3,1,2) and / or Cast(to FLOAT) on its input.  Both are
  // no-ops here: the engine's activation layout is its own (NHWC fp16) and the pre-kernel produces the values directly.
  std::set<std::string> input_chain = {in0.name};
  for (auto& n : g.nodes) {
    if (n.op == "Identity") { if (n.in.size() == 1 && n.out.size() == 1) alias[n.out[0]] = n.in[0]; continue; }
    if ((n.op == "Cast" || n.op == "Transpose") && n.in.size() == 1 && n.out.size() == 1 && input_chain.count(resolve(n.in[0])) &&
        !(n.op == "Cast" && shape_values.count(n.in[0]))) {
      if (n.op == "Cast" && attr_i(n, "to", 0) != 1) throw ModelError(INFUR_E_MODEL_LOAD, "Cast '" + n.name + "' on the model input must convert to FLOAT");
      if (n.op == "Transpose") {
        auto perm = attr_ints(n, "perm", {});
        if (m.io.nchw || perm != std::vector<int64_t>{0, 3, 1, 2}) throw ModelError(INFUR_E_MODEL_LOAD, "Transpose '" + n.name + "' on the model input must be NHWC -> NCHW (perm 0,3,1,2)");
      }
      alias[n.out[0]] = n.in[0];
      input_chain.insert(n.out[0]);
      continue;
    }
    if (n.op == "Conv" || n.op == "Relu" || n.op == "Add" || n.op == "MaxPool" || n.op == "Resize" || n.op == "QLinearConv" ||
        n.op == "QLinearAdd" || n.op == "QuantizeLinear" || n.op == "DequantizeLinear") { compute.push_back(&n); continue; }
    if (shape_ops.count(n.op)) {
      // only allowed when it does not touch an activation except through Shape
      for (auto& o : n.out) shape_values.insert(o);
      continue;
    }
    throw ModelError(INFUR_E_MODEL_LOAD, "unsupported operator '" + n.op + "' (node '" + n.name + "'); supported: Conv, Relu, Add, MaxPool, Resize and the "
                     "QOperator forms QuantizeLinear, QLinearConv, QLinearAdd, DequantizeLinear");
  }
  std::map<std::string, int> consumers;
  std::map<std::string, const OnnxNode*> single_consumer;
  for (auto* n : compute) {
    // activation inputs only (weights, scales, zero points and sizes are not tensors of the lowered graph)
    std::vector<size_t> data;
    if (n->op == "QLinearAdd") data = {0, 3};
    else if (n->op == "Add") data = {0, 1};
    else data = {0};
    for (size_t i : data) {
      if (i >= n->in.size()) continue;
      std::string r = resolve(n->in[i]);
      consumers[r]++;
      single_consumer[r] = n;
    }
  }
  for (auto& vi : g.outputs) consumers[resolve(vi.name)] += 1000;  // graph outputs are never fused away
  // a Shape node reading an activation is a consumer too, but a harmless one (sizes only)

  // ---- pass 2: emit fused ops
  std::map<std::string, int> tid;  // tensor name -> id
  auto new_tensor = [&](const std::string& name, int channels) {
    int id = m.num_tensors++;
    tid[name] = id;
    m.tensor_channels.push_back(channels);
    return id;
  };
  m.input_tensor = new_tensor(in0.name, 3);
  auto need = [&](const std::string& name, const OnnxNode& n) {
    auto it = tid.find(resolve(name));
    if (it == tid.end()) throw ModelError(INFUR_E_MODEL_LOAD, "node '" + n.name + "' (" + n.op + ") reads '" + name + "' which no supported node produces");
    return it->second;
  };
  std::set<const OnnxNode*> fused;
  struct Pending { LoweredOp op; std::string in_name; };
  std::map<std::string, LoweredOp> pending;  // conv outputs waiting for their Add
  auto emit = [&](LoweredOp&& op, const std::string& out_name) {
    op.out = new_tensor(out_name, op.kind == OpKind::Conv ? op.conv.cout : m.tensor_channels[op.in]);
    m.ops.push_back(std::move(op));
  };
  auto take_relu = [&](const std::string& out_name, std::string& final_name) -> bool {
    if (consumers[out_name] == 1 && single_consumer[out_name]->op == "Relu") {
      const OnnxNode* r = single_consumer[out_name];
      fused.insert(r);
      final_name = r->out[0];
      return true;
    }
    final_name = out_name;
    return false;
  };

  // ---- quantised (QOperator) graphs: per-tensor (scale, zero point) of every quantised activation
  struct QInfo { float scale; int zp, qmin, qmax; };
  std::map<int, QInfo> qinfo;          // tensor id -> quantisation; absent = float tensor
  std::map<std::string, QInfo> pending_q;  // output quantisation of a pending QLinearConv
  std::set<int> dequantized;           // tensor ids whose stored form is the de-quantised f32 value (head inputs of Resize)
  auto q_init = [&](const OnnxNode& n, size_t idx, const char* what) -> const OnnxTensor& {
    if (idx >= n.in.size() || n.in[idx].empty()) throw ModelError(INFUR_E_MODEL_LOAD, n.op + " '" + n.name + "': missing " + what);
    auto it = g.inits.find(resolve(n.in[idx]));
    if (it == g.inits.end()) throw ModelError(INFUR_E_MODEL_LOAD, n.op + " '" + n.name + "': " + what + " is not an initializer");
    return it->second;
  };
  auto q_scale = [&](const OnnxNode& n, size_t idx, const char* what) {
    std::vector<float> v = tensor_f32(q_init(n, idx, what), n.op + " '" + n.name + "'");
    if (v.size() != 1 || !(v[0] > 0.f)) throw ModelError(INFUR_E_MODEL_LOAD, n.op + " '" + n.name + "': " + what + " must be one positive FLOAT");
    return v[0];
  };
  auto q_zp = [&](const OnnxNode& n, size_t idx, const char* what, float scale) {
    QInfo q{scale, 0, 0, 255};
    if (idx >= n.in.size() || n.in[idx].empty()) return q;   // optional: uint8 zero
    const OnnxTensor& t = q_init(n, idx, what);
    std::vector<int32_t> v = tensor_i32(t, n.op + " '" + n.name + "'");
    if (v.size() != 1 || (t.dtype != 2 && t.dtype != 3)) throw ModelError(INFUR_E_MODEL_LOAD, n.op + " '" + n.name + "': " + what + " must be one UINT8 / INT8 value");
    q.zp = v[0];
    if (t.dtype == 3) { q.qmin = -128; q.qmax = 127; }
    return q;
  };
  auto same_q = [](const QInfo& a, const QInfo& b) { return a.scale == b.scale && a.zp == b.zp && a.qmin == b.qmin; };
  auto need_q = [&](int tensor, const QInfo& want, const OnnxNode& n, const char* which) {
    auto it = qinfo.find(tensor);
    if (it == qinfo.end()) throw ModelError(INFUR_E_MODEL_LOAD, n.op + " '" + n.name + "': input " + which + " is not a quantised tensor");
    if (!same_q(it->second, want))
      throw ModelError(INFUR_E_MODEL_LOAD, n.op + " '" + n.name + "': scale / zero point of input " + which + " differ from its producer's (re-quantising edges are not supported)");
  };
  auto need_float = [&](int tensor, const OnnxNode& n) {
    if (qinfo.count(tensor)) throw ModelError(INFUR_E_MODEL_LOAD, n.op + " '" + n.name + "' reads a quantised tensor (mixed float / quantised graphs are not supported)");
  };
  auto conv_geometry = [&](const OnnxNode& n, const OnnxTensor& wt, ConvOp& c) {
    if (wt.dims.size() != 4) throw ModelError(INFUR_E_MODEL_LOAD, n.op + " '" + n.name + "': weight must be 4-D");
    for (int64_t d : wt.dims) if (d < 1 || d > (1 << 20)) throw ModelError(INFUR_E_MODEL_LOAD, n.op + " '" + n.name + "': weight dimensions must be 1..2^20");
    c.cout = (int)wt.dims[0]; c.cin = (int)wt.dims[1]; c.kh = (int)wt.dims[2]; c.kw = (int)wt.dims[3];
    if (attr_i(n, "group", 1) != 1) throw ModelError(INFUR_E_MODEL_LOAD, n.op + " '" + n.name + "': group != 1 is not supported");
    auto strides = attr_ints(n, "strides", {1, 1}), dil = attr_ints(n, "dilations", {1, 1}), pads = attr_ints(n, "pads", {0, 0, 0, 0});
    auto ks = attr_ints(n, "kernel_shape", {c.kh, c.kw});
    if (ks.size() != 2 || ks[0] != c.kh || ks[1] != c.kw) throw ModelError(INFUR_E_MODEL_LOAD, n.op + " '" + n.name + "': kernel_shape does not match the weight");
    if (auto* ap = n.attr("auto_pad")) if (!ap->s.empty() && ap->s != "NOTSET") throw ModelError(INFUR_E_MODEL_LOAD, n.op + " '" + n.name + "': auto_pad is not supported");
    if (strides.size() != 2 || strides[0] != strides[1] || dil.size() != 2 || dil[0] != dil[1] || pads.size() != 4 || !all_eq(pads, pads[0]) || c.kh != c.kw)
      throw ModelError(INFUR_E_MODEL_LOAD, n.op + " '" + n.name + "': only square kernels with symmetric stride/dilation/padding are supported");
    check_geometry(n.op + " '" + n.name + "'", c.kh, strides[0], dil[0], pads[0]);
    c.stride = (int)strides[0]; c.dil = (int)dil[0]; c.pad = (int)pads[0];
  };

  // ONNX defines Relu on float (and signed integer) tensors only; quantisers fold a ReLU into the u8 clamp of the producing node
  // (zero point = lowest code).  A graph that applies Relu to a quantised tensor is rejected rather than guessed at.
  auto no_relu_on_quantised = [&](const std::string& out_name) {
    auto it = single_consumer.find(out_name);
    if (it != single_consumer.end() && it->second->op == "Relu")
      throw ModelError(INFUR_E_MODEL_LOAD, "Relu '" + it->second->name + "' reads a quantised tensor: not valid ONNX (a quantiser folds ReLU into the clamp of the producing node)");
  };

  for (auto* np : compute) {
    const OnnxNode& n = *np;
    if (fused.count(np)) continue;
    if (n.op == "QuantizeLinear") {
      // only on the network input: the pre-kernel's lookup table produces the quantised values directly
      if (n.in.empty() || !input_chain.count(resolve(n.in[0])))
        throw ModelError(INFUR_E_MODEL_LOAD, "QuantizeLinear '" + n.name + "': only supported on the network input");
      const float sc = q_scale(n, 1, "y_scale");
      QInfo q = q_zp(n, 2, "y_zero_point", sc);
      m.quant = true; m.in_scale = sc; m.in_zp = q.zp; m.in_qmin = q.qmin; m.in_qmax = q.qmax;
      tid[n.out[0]] = m.input_tensor;
      qinfo[m.input_tensor] = q;
    } else if (n.op == "QLinearConv") {
      // x, x_scale, x_zero_point, w, w_scale, w_zero_point, y_scale, y_zero_point, [B]
      if (n.in.size() < 8) throw ModelError(INFUR_E_MODEL_LOAD, "QLinearConv '" + n.name + "' needs 8 or 9 inputs");
      const OnnxTensor& wt = q_init(n, 3, "weight");
      LoweredOp op; op.kind = OpKind::Conv; op.name = n.name.empty() ? n.out[0] : n.name;
      ConvOp& c = op.conv;
      conv_geometry(n, wt, c);
      c.quant = true;
      const float xs = q_scale(n, 1, "x_scale");
      const QInfo xq = q_zp(n, 2, "x_zero_point", xs);
      const float ys = q_scale(n, 6, "y_scale");
      const QInfo yq = q_zp(n, 7, "y_zero_point", ys);
      std::vector<float> ws = tensor_f32(q_init(n, 4, "w_scale"), "QLinearConv '" + n.name + "'");
      std::vector<int32_t> wz = tensor_i32(q_init(n, 5, "w_zero_point"), "QLinearConv '" + n.name + "'");
      if ((ws.size() != 1 && (int)ws.size() != c.cout) || (wz.size() != 1 && (int)wz.size() != c.cout))
        throw ModelError(INFUR_E_MODEL_LOAD, "QLinearConv '" + n.name + "': w_scale / w_zero_point must be scalars or per-output-channel");
      std::vector<int32_t> wq = tensor_i32(wt, "QLinearConv '" + n.name + "'");
      if (wt.dtype == 6) throw ModelError(INFUR_E_MODEL_LOAD, "QLinearConv '" + n.name + "': weight must be UINT8 or INT8");
      c.weight.resize(wq.size());
      for (int o = 0; o < c.cout; ++o) {
        const int z = wz[wz.size() == 1 ? 0 : o];
        for (int i = 0; i < c.cin; ++i)
          for (int y = 0; y < c.kh; ++y)
            for (int x = 0; x < c.kw; ++x)
              c.weight[(((size_t)o * c.kh + y) * c.kw + x) * c.cin + i] = (float)(wq[(((size_t)o * c.cin + i) * c.kh + y) * c.kw + x] - z);
      }
      c.bias.assign(c.cout, 0.f);
      if (n.in.size() >= 9 && !n.in[8].empty()) {
        const OnnxTensor& bt = q_init(n, 8, "bias");
        if (bt.dtype != 6) throw ModelError(INFUR_E_MODEL_LOAD, "QLinearConv '" + n.name + "': bias must be INT32");
        std::vector<int32_t> b = tensor_i32(bt, "QLinearConv '" + n.name + "'");
        if ((int)b.size() != c.cout) throw ModelError(INFUR_E_MODEL_LOAD, "QLinearConv '" + n.name + "': bias length mismatch");
        for (int o = 0; o < c.cout; ++o) {
          if (b[o] >= (1 << 24) || b[o] <= -(1 << 24))
            throw ModelError(INFUR_E_MODEL_LOAD, "QLinearConv '" + n.name + "': |bias| >= 2^24 cannot be carried exactly by the f32 accumulator");
          c.bias[o] = (float)b[o];
        }
      }
      c.qmul.resize(c.cout);
      for (int o = 0; o < c.cout; ++o) {
        // ONNX Runtime's QLinearConv: output_scale[c] = (x_scale * w_scale[c]) / y_scale, evaluated in f32
        volatile float t = xs * ws[ws.size() == 1 ? 0 : o];
        volatile float q = t / ys;
        c.qmul[o] = q;
      }
      c.q_lo = (float)(yq.qmin - yq.zp); c.q_hi = (float)(yq.qmax - yq.zp);
      c.x_zp = xq.zp; c.out_zp = yq.zp; c.all_u8 = xq.qmin == 0 && yq.qmin == 0;
      op.in = need(n.in[0], n);
      need_q(op.in, xq, n, "x");
      if (m.tensor_channels[op.in] != c.cin) throw ModelError(INFUR_E_MODEL_LOAD, "QLinearConv '" + n.name + "': input has " + std::to_string(m.tensor_channels[op.in]) + " channels, weight expects " + std::to_string(c.cin));
      const std::string out_name = n.out[0];
      if (consumers[out_name] == 1 && single_consumer[out_name]->op == "QLinearAdd") { pending[out_name] = std::move(op); pending_q[out_name] = yq; continue; }
      no_relu_on_quantised(out_name);
      emit(std::move(op), out_name);
      qinfo[tid[out_name]] = yq;
    } else if (n.op == "QLinearAdd") {
      // A, A_scale, A_zero_point, B, B_scale, B_zero_point, C_scale, C_zero_point   (com.microsoft)
      if (n.in.size() < 7) throw ModelError(INFUR_E_MODEL_LOAD, "QLinearAdd '" + n.name + "' needs 7 or 8 inputs");
      const std::string a = resolve(n.in[0]), b = resolve(n.in[3]);
      const bool a_main = pending.count(a) != 0;
      if (!a_main && !pending.count(b)) throw ModelError(INFUR_E_MODEL_LOAD, "QLinearAdd '" + n.name + "': neither input is a convolution output (stand-alone QLinearAdd is not supported)");
      const std::string main_name = a_main ? a : b, other = a_main ? b : a;
      const size_t mi = a_main ? 0 : 3, oi = a_main ? 3 : 0;
      const float ms = q_scale(n, mi + 1, "scale"), os_ = q_scale(n, oi + 1, "scale"), cs = q_scale(n, 6, "C_scale");
      const QInfo mq = q_zp(n, mi + 2, "zero point", ms), oq = q_zp(n, oi + 2, "zero point", os_), cq = q_zp(n, 7, "C_zero_point", cs);
      if (pending.count(other)) {  // the projection shortcut of a bottleneck: runs on its own first
        LoweredOp o = std::move(pending[other]); pending.erase(other);
        emit(std::move(o), other);
        qinfo[tid[other]] = pending_q[other];
      }
      LoweredOp op = std::move(pending[main_name]); pending.erase(main_name);
      if (!same_q(pending_q[main_name], mq)) throw ModelError(INFUR_E_MODEL_LOAD, "QLinearAdd '" + n.name + "': scale / zero point of the convolution input differ from the QLinearConv's output");
      op.conv.residual = need(other, n);
      need_q(op.conv.residual, oq, n, "(residual)");
      if (m.tensor_channels[op.conv.residual] != op.conv.cout) throw ModelError(INFUR_E_MODEL_LOAD, "QLinearAdd '" + n.name + "': channel mismatch");
      // ONNX Runtime's QLinearAdd: C = sat(rne(A_scale / C_scale * (A - A_zp) + B_scale / C_scale * (B - B_zp)) + C_zp), f32
      { volatile float ra = ms / cs, rb = os_ / cs; op.conv.q_ra = ra; op.conv.q_rb = rb; }
      op.conv.q_lo2 = (float)(cq.qmin - cq.zp); op.conv.q_hi2 = (float)(cq.qmax - cq.zp);
      op.conv.out_zp = cq.zp; op.conv.res_zp = oq.zp; op.conv.all_u8 = op.conv.all_u8 && cq.qmin == 0 && oq.qmin == 0;
      no_relu_on_quantised(n.out[0]);
      emit(std::move(op), n.out[0]);
      qinfo[tid[n.out[0]]] = cq;
    } else if (n.op == "DequantizeLinear") {
      // only as the step between a head convolution and its Resize: the convolution stores the de-quantised f32 logits
      const int t = need(n.in[0], n);
      const float sc = q_scale(n, 1, "x_scale");
      need_q(t, q_zp(n, 2, "x_zero_point", sc), n, "x");
      int prod = -1;
      for (size_t j = 0; j < m.ops.size(); ++j) if (m.ops[j].out == t) prod = (int)j;
      if (prod < 0 || m.ops[prod].kind != OpKind::Conv || m.ops[prod].conv.residual >= 0 || consumers[resolve(n.in[0])] != 1)
        throw ModelError(INFUR_E_MODEL_LOAD, "DequantizeLinear '" + n.name + "': only supported on a convolution output that nothing else reads (the logits before Resize)");
      m.ops[prod].conv.deq_scale = sc;
      tid[n.out[0]] = t;
      dequantized.insert(t);
    } else if (n.op == "Conv") {
      if (n.in.size() < 2) throw ModelError(INFUR_E_MODEL_LOAD, "Conv '" + n.name + "' has no weight input");
      auto wit = g.inits.find(resolve(n.in[1]));
      if (wit == g.inits.end()) throw ModelError(INFUR_E_MODEL_LOAD, "Conv '" + n.name + "': weight is not an initializer (quantised / dynamic weights are not supported)");
      const OnnxTensor& wt = wit->second;
      if (wt.dims.size() != 4) throw ModelError(INFUR_E_MODEL_LOAD, "Conv '" + n.name + "': weight must be 4-D");
      for (int64_t d : wt.dims) if (d < 1 || d > (1 << 20)) throw ModelError(INFUR_E_MODEL_LOAD, "Conv '" + n.name + "': weight dimensions must be 1..2^20");
      LoweredOp op; op.kind = OpKind::Conv; op.name = n.name.empty() ? n.out[0] : n.name;
      ConvOp& c = op.conv;
      c.cout = (int)wt.dims[0]; c.cin = (int)wt.dims[1]; c.kh = (int)wt.dims[2]; c.kw = (int)wt.dims[3];
      if (attr_i(n, "group", 1) != 1) throw ModelError(INFUR_E_MODEL_LOAD, "Conv '" + n.name + "': group != 1 is not supported");
      auto strides = attr_ints(n, "strides", {1, 1}), dil = attr_ints(n, "dilations", {1, 1}), pads = attr_ints(n, "pads", {0, 0, 0, 0});
      auto ks = attr_ints(n, "kernel_shape", {c.kh, c.kw});
      if (ks.size() != 2 || ks[0] != c.kh || ks[1] != c.kw) throw ModelError(INFUR_E_MODEL_LOAD, "Conv '" + n.name + "': kernel_shape does not match the weight");
      if (auto* ap = n.attr("auto_pad")) if (!ap->s.empty() && ap->s != "NOTSET") throw ModelError(INFUR_E_MODEL_LOAD, "Conv '" + n.name + "': auto_pad is not supported");
      if (strides.size() != 2 || strides[0] != strides[1] || dil.size() != 2 || dil[0] != dil[1] || pads.size() != 4 || !all_eq(pads, pads[0]) || c.kh != c.kw)
        throw ModelError(INFUR_E_MODEL_LOAD, "Conv '" + n.name + "': only square kernels with symmetric stride/dilation/padding are supported");
      check_geometry("Conv '" + n.name + "'", c.kh, strides[0], dil[0], pads[0]);
      c.stride = (int)strides[0]; c.dil = (int)dil[0]; c.pad = (int)pads[0];
      std::vector<float> w = tensor_f32(wt, "Conv '" + n.name + "'");
      c.weight.resize(w.size());
      for (int o = 0; o < c.cout; ++o)
        for (int i = 0; i < c.cin; ++i)
          for (int y = 0; y < c.kh; ++y)
            for (int x = 0; x < c.kw; ++x)
              c.weight[(((size_t)o * c.kh + y) * c.kw + x) * c.cin + i] = w[(((size_t)o * c.cin + i) * c.kh + y) * c.kw + x];
      if (n.in.size() >= 3 && !n.in[2].empty()) {
        auto bit = g.inits.find(resolve(n.in[2]));
        if (bit == g.inits.end()) throw ModelError(INFUR_E_MODEL_LOAD, "Conv '" + n.name + "': bias is not an initializer");
        c.bias = tensor_f32(bit->second, "Conv '" + n.name + "'");
        if ((int)c.bias.size() != c.cout) throw ModelError(INFUR_E_MODEL_LOAD, "Conv '" + n.name + "': bias length mismatch");
      } else c.bias.assign(c.cout, 0.f);
      op.in = need(n.in[0], n);
      need_float(op.in, n);
      if (m.tensor_channels[op.in] != c.cin) throw ModelError(INFUR_E_MODEL_LOAD, "Conv '" + n.name + "': input has " + std::to_string(m.tensor_channels[op.in]) + " channels, weight expects " + std::to_string(c.cin));
      const std::string out_name = n.out[0];
      if (consumers[out_name] == 1 && single_consumer[out_name]->op == "Add") { pending[out_name] = std::move(op); continue; }
      std::string final_name;
      op.conv.relu = take_relu(out_name, final_name);
      emit(std::move(op), final_name);
    } else if (n.op == "Add") {
      if (n.in.size() != 2) throw ModelError(INFUR_E_MODEL_LOAD, "Add '" + n.name + "' must have two inputs");
      std::string a = resolve(n.in[0]), b = resolve(n.in[1]);
      std::string main_name, other;
      if (pending.count(a)) { main_name = a; other = b; }
      else if (pending.count(b)) { main_name = b; other = a; }
      else throw ModelError(INFUR_E_MODEL_LOAD, "Add '" + n.name + "': neither input is a convolution output (stand-alone Add is not supported)");
      if (pending.count(other)) {  // e.g. the downsample conv of a bottleneck: runs on its own first
        LoweredOp o = std::move(pending[other]); pending.erase(other);
        emit(std::move(o), other);
      }
      LoweredOp op = std::move(pending[main_name]); pending.erase(main_name);
      op.conv.residual = need(other, n);
      need_float(op.conv.residual, n);
      if (m.tensor_channels[op.conv.residual] != op.conv.cout) throw ModelError(INFUR_E_MODEL_LOAD, "Add '" + n.name + "': channel mismatch");
      std::string final_name;
      op.conv.relu = take_relu(n.out[0], final_name);
      emit(std::move(op), final_name);
    } else if (n.op == "Relu") {
      throw ModelError(INFUR_E_MODEL_LOAD, "Relu '" + n.name + "' does not follow a Conv or Conv+Add (stand-alone Relu is not supported)");
    } else if (n.op == "MaxPool") {
      auto ks = attr_ints(n, "kernel_shape", {}), st = attr_ints(n, "strides", {1, 1}), pads = attr_ints(n, "pads", {0, 0, 0, 0}), dl = attr_ints(n, "dilations", {1, 1});
      if (ks.size() != 2 || ks[0] != ks[1] || st.size() != 2 || st[0] != st[1] || pads.size() != 4 || !all_eq(pads, pads[0]) || attr_i(n, "ceil_mode", 0) != 0 || !all_eq(dl, 1))
        throw ModelError(INFUR_E_MODEL_LOAD, "MaxPool '" + n.name + "': only square, symmetric, floor-mode pooling is supported");
      check_geometry("MaxPool '" + n.name + "'", ks[0], st[0], 1, pads[0]);
      LoweredOp op; op.kind = OpKind::MaxPool; op.name = n.name.empty() ? n.out[0] : n.name;
      op.pool_k = (int)ks[0]; op.pool_s = (int)st[0]; op.pool_p = (int)pads[0];
      op.in = need(n.in[0], n);
      const int pool_in = op.in;
      emit(std::move(op), n.out[0]);
      if (qinfo.count(pool_in)) qinfo[tid[n.out[0]]] = qinfo[pool_in];   // max commutes with the affine quantisation map
    } else if (n.op == "Resize") {
      std::string mode = n.attr("mode") ? n.attr("mode")->s : "nearest";
      std::string ctm = n.attr("coordinate_transformation_mode") ? n.attr("coordinate_transformation_mode")->s : "half_pixel";
      if (mode != "linear" || (ctm != "half_pixel" && ctm != "pytorch_half_pixel"))
        throw ModelError(INFUR_E_MODEL_LOAD, "Resize '" + n.name + "': only mode=linear with half_pixel / pytorch_half_pixel is supported (got " + mode + ", " + ctm + ")");
      // the target size must come from the size-computing subgraph (Shape of the network input), i.e.
      // "resize to the input's H x W" -- the only form FCN uses.
      bool sized = n.in.size() >= 4 && shape_values.count(n.in[3]);
      if (!sized) throw ModelError(INFUR_E_MODEL_LOAD, "Resize '" + n.name + "': expected a computed `sizes` input (resize to the network input size)");
      bool is_output = false;
      for (auto& vi : g.outputs) if (vi.name == n.out[0]) is_output = true;
      if (!is_output) throw ModelError(INFUR_E_MODEL_LOAD, "Resize '" + n.name + "': only supported as the last node of an output head");
      LoweredHead hd; hd.name = n.out[0]; hd.tensor = need(n.in[0], n); hd.num_classes = m.tensor_channels[hd.tensor];
      if (qinfo.count(hd.tensor) && !dequantized.count(hd.tensor))
        throw ModelError(INFUR_E_MODEL_LOAD, "Resize '" + n.name + "' reads a quantised tensor; expected DequantizeLinear before it");
      m.heads.push_back(hd);
    }
  }
  if (!pending.empty()) throw ModelError(INFUR_E_MODEL_LOAD, "internal: convolution output left without its Add");
  for (auto& op : m.ops)
    if (op.kind == OpKind::Conv && op.conv.quant != m.quant)
      throw ModelError(INFUR_E_MODEL_LOAD, "convolution '" + op.name + "': float and quantised convolutions cannot be mixed in one model");
  if (m.quant)
    for (auto& h : m.heads)
      if (!dequantized.count(h.tensor)) throw ModelError(INFUR_E_MODEL_LOAD, "head '" + h.name + "' of a quantised model is not fed by DequantizeLinear");
  if (m.heads.empty()) throw ModelError(INFUR_E_MODEL_LOAD, "model has no `Resize` output head (not an FCN-style segmentation model)");
  // heads in graph-output order
  std::vector<LoweredHead> ordered;
  for (auto& vi : g.outputs) for (auto& h : m.heads) if (h.name == vi.name) ordered.push_back(h);
  m.heads = ordered;
}

int fuse_projection_shortcuts(LoweredModel& m) {
  int fused = 0;
  for (size_t i = 0; i < m.ops.size(); ++i) {
    LoweredOp& op = m.ops[i];
    if (op.kind != OpKind::Conv) continue;
    ConvOp& c = op.conv;
    if (c.quant) continue;   // each QLinearConv re-quantises its own output: the two sums cannot share an accumulator
    if (c.residual < 0 || c.in2 >= 0 || c.kh != 1 || c.kw != 1 || c.stride != 1 || c.pad != 0 || c.cin % 64 != 0) continue;
    // producer of the residual: a bare 1x1 convolution whose output nobody else reads
    int pj = -1, uses = 0;
    for (size_t j = 0; j < m.ops.size(); ++j) {
      if (m.ops[j].out == c.residual) pj = (int)j;
      if (m.ops[j].in == c.residual) ++uses;
      if (m.ops[j].kind == OpKind::Conv && (m.ops[j].conv.residual == c.residual || m.ops[j].conv.in2 == c.residual)) ++uses;
    }
    for (auto& h : m.heads) if (h.tensor == c.residual) ++uses;
    if (pj < 0 || pj >= (int)i || uses != 1 || m.ops[pj].kind != OpKind::Conv) continue;
    const ConvOp& d = m.ops[pj].conv;
    if (d.kh != 1 || d.kw != 1 || d.pad != 0 || d.relu || d.residual >= 0 || d.in2 >= 0 || d.cout != c.cout || d.cin % 64 != 0 ||
        (d.stride != 1 && d.stride != 2))
      continue;
    c.in2 = m.ops[pj].in; c.cin2 = d.cin; c.stride2 = d.stride; c.weight2 = d.weight;
    for (int co = 0; co < c.cout; ++co) c.bias[co] += d.bias[co];
    c.residual = -1;
    op.name += " + " + m.ops[pj].name;
    m.ops.erase(m.ops.begin() + pj);
    --i;
    ++fused;
  }
  return fused;
}

bool int8_plan_eligible(const LoweredModel& m, const std::vector<char>* needed, std::string* why) {
  auto no = [&](const std::string& reason) { if (why) *why = reason; return false; };
  if (!m.quant) return no("not a quantised model");
  std::vector<char> is_head(m.num_tensors, 0);
  for (auto& h : m.heads) is_head[h.tensor] = 1;
  for (size_t i = 0; i < m.ops.size(); ++i) {
    if (needed && !(*needed)[i]) continue;
    const LoweredOp& op = m.ops[i];
    if (op.kind == OpKind::MaxPool) {
      if (!(op.pool_k == 3 && op.pool_s == 2 && op.pool_p == 1)) return no("MaxPool '" + op.name + "' is not 3x3 / stride 2 / pad 1");
      continue;
    }
    const ConvOp& c = op.conv;
    const bool stem = op.in == m.input_tensor;
    if (!c.all_u8) return no("convolution '" + op.name + "' touches a tensor that is not uint8");
    if (stem && !(c.cin == 3 && c.kh == 7 && c.stride == 2 && c.pad == 3)) return no("the convolution on the network input, '" + op.name + "', is not the 7x7 / stride 2 RGB stem");
    if (!stem && c.x_zp != 0) return no("input of convolution '" + op.name + "' has zero point " + std::to_string(c.x_zp) + " (zero padding would not be the padding value)");
    if (is_head[op.out] && c.residual >= 0) return no("head convolution '" + op.name + "' carries a residual");
    if (!stem)
      for (float w : c.weight)
        if (w < -128.f || w > 127.f) return no("weights of convolution '" + op.name + "' do not fit int8 after subtracting their zero point");
  }
  return true;
}

std::string describe(const LoweredModel& m) {
  std::ostringstream os;
  os << "inputs:";
  for (auto& s : m.io.input_names) os << " " << s;
  os << " dtype=" << m.io.input0_dtype << " layout=" << (m.io.nchw ? "NCHW" : "NHWC") << " color=" << (m.io.rgb ? "RGB" : "BGR");
  if (m.quant) os << " quantised(zp=" << m.in_zp << ")";
  os << "\n";
  os << "outputs:";
  for (auto& s : m.io.output_names) os << " " << s;
  os << "\n";
  for (size_t i = 0; i < m.ops.size(); ++i) {
    const LoweredOp& o = m.ops[i];
    if (o.kind == OpKind::Conv) {
      const ConvOp& c = o.conv;
      os << i << " conv t" << o.in << "->t" << o.out << " " << c.cin << "->" << c.cout << " k" << c.kh << " s" << c.stride << " p" << c.pad << " d" << c.dil
         << (c.residual >= 0 ? " +t" + std::to_string(c.residual) : std::string()) << (c.relu ? " relu" : "");
      if (c.quant) {
        os << " q[" << c.q_lo << "," << c.q_hi << "]";
        if (c.residual >= 0) os << " add[" << c.q_lo2 << "," << c.q_hi2 << "]";
        if (c.deq_scale != 0.f) os << " deq";
      }
      os << "\n";
    } else {
      os << i << " maxpool t" << o.in << "->t" << o.out << " k" << o.pool_k << " s" << o.pool_s << " p" << o.pool_p << "\n";
    }
  }
  for (auto& h : m.heads) os << "head " << h.name << " t" << h.tensor << " classes=" << h.num_classes << "\n";
  if (m.quant) {
    std::string why;
    if (int8_plan_eligible(m, nullptr, &why)) os << "plan: int8 (u8 tensors, tcgen05.mma.kind::i8)\n";
    else os << "plan: fp16-carried integers (no int8 plan: " << why << ")\n";
  }
  return os.str();
}

}  // namespace infur
