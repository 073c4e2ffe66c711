// Minimal ONNX (protobuf wire format) reader and graph lowering for the FCN-ResNet family.
//
// Replaces what `with_model_from_file` does for the reference (infur/src/predict_onnx.rs:288-293:
// ONNX Runtime parses the file and optimises the graph at GraphOptimizationLevel::Extended).  No
// protobuf library: the handful of message fields needed are decoded straight from the wire format.
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace infur {

struct OnnxTensor {
  std::string name;
  std::vector<int64_t> dims;
  int32_t dtype = 0;            // TensorProto.DataType: 1 = FLOAT, 7 = INT64, ...
  const uint8_t* raw = nullptr; // raw_data (points into the file buffer)
  size_t raw_size = 0;
  std::vector<float> float_data;
  std::vector<int64_t> int64_data;
  std::vector<int32_t> int32_data;   // also carries INT8 / UINT8 / INT32 values when raw_data is not used
  size_t numel() const { size_t n = 1; for (auto d : dims) n *= (size_t)d; return n; }
};

struct OnnxAttr {
  std::string name;
  int64_t i = 0;
  float f = 0.f;
  std::string s;
  std::vector<int64_t> ints;
  bool has_t = false;
  OnnxTensor t;
};

struct OnnxNode {
  std::string op, name;
  std::vector<std::string> in, out;
  std::vector<OnnxAttr> attrs;
  const OnnxAttr* attr(const char* n) const {
    for (auto& a : attrs) if (a.name == n) return &a;
    return nullptr;
  }
};

struct OnnxValueInfo {
  std::string name;
  int32_t elem_type = 0;
  bool has_shape = false;
  std::vector<int64_t> dims;  // -1 for symbolic / unknown
};

struct OnnxGraph {
  std::vector<uint8_t> file;  // owns the bytes raw pointers refer to
  int64_t ir_version = 0, opset = 0;
  std::string producer;
  std::vector<OnnxNode> nodes;
  std::map<std::string, OnnxTensor> inits;
  std::vector<OnnxValueInfo> inputs, outputs;
};

// Throws std::runtime_error with a readable message on malformed input.
void parse_onnx(std::vector<uint8_t>&& bytes, OnnxGraph& g);
void read_file(const std::string& path, std::vector<uint8_t>& bytes);

// ---------------------------------------------------------------------------------------------
// Lowered model: the fused operator list the engine executes.

enum class OpKind { Conv, MaxPool };

struct ConvOp {
  int cin = 0, cout = 0, kh = 1, kw = 1, stride = 1, pad = 0, dil = 1;
  bool relu = false;
  int residual = -1;                 // tensor id added before the ReLU, or -1
  std::vector<float> weight;         // [cout][kh][kw][cin]  (OHWI, f32; packed to fp16 at upload)
  std::vector<float> bias;           // [cout]
  // Fused projection shortcut (fuse_projection_shortcuts): a second 1x1 / pad 0 convolution with stride `stride2`
  // over tensor `in2` accumulated into the same output: out = act(conv(in) + conv2(in2) + bias), bias = b + b2.
  int in2 = -1, cin2 = 0, stride2 = 1;
  std::vector<float> weight2;        // [cout][cin2]
  // Quantised convolution (QLinearConv, optionally followed by a QLinearAdd).  Activations are carried as the integers
  // q - zero_point (|.| <= 255, exact in fp16), `weight` holds w_q - w_zero_point and `bias` the int32 bias, so the fp16
  // tensor-core product with f32 accumulation IS the integer accumulator while it stays below 2^24.  The epilogue
  // requantises like ONNX Runtime's QLinearConv / QLinearAdd:
  //   r = clamp(rne((acc + bias) * qmul[c]), q_lo, q_hi)                 qmul[c] = (x_scale * w_scale[c]) / y_scale
  //   r = clamp(rne(r * q_ra + res * q_rb), q_lo2, q_hi2)                (only with `residual`; q_ra = a_scale / c_scale ...)
  // and a head convolution followed by DequantizeLinear stores r * deq_scale as f32 logits.
  bool quant = false;
  std::vector<float> qmul;           // [cout]
  float q_lo = 0.f, q_hi = 0.f, q_ra = 0.f, q_rb = 0.f, q_lo2 = 0.f, q_hi2 = 0.f;
  float deq_scale = 0.f;
  // zero points of the input, of the stored output (the Add's when there is one) and of the residual, and whether all of
  // those tensors are u8: what an int8 plan (u8 activation tensors holding the raw q) needs to know
  int x_zp = 0, out_zp = 0, res_zp = 0;
  bool all_u8 = true;
};

struct LoweredOp {
  OpKind kind = OpKind::Conv;
  int in = -1, out = -1;             // tensor ids
  ConvOp conv;                       // kind == Conv
  int pool_k = 3, pool_s = 2, pool_p = 1;  // kind == MaxPool
  std::string name;
};

struct LoweredHead {
  std::string name;   // graph output name ("out", "aux")
  int tensor = -1;    // low-res logits tensor id feeding the final Resize
  int num_classes = 0;
};

// Input conventions inferred exactly like infer_img_pre_proc (predict_onnx.rs:223-265) and
// Model::control (:296-306).
struct ModelIO {
  std::vector<std::string> input_names, output_names;
  std::string input0_dtype;   // "Float" / "Uint8" (Debug print of TensorElementDataType)
  bool nchw = true;           // DimSeq
  bool rgb = true;            // ColorSeq
  bool float_input = true;    // ColorRange::Float32(norm)
};

struct LoweredModel {
  ModelIO io;
  int num_tensors = 0;
  int input_tensor = 0;
  std::vector<LoweredOp> ops;        // topological order
  std::vector<LoweredHead> heads;    // in graph-output order
  std::vector<int> tensor_channels;  // per tensor id
  // QOperator-format (quantised) model: QuantizeLinear on the network input, q = sat(rne(x / in_scale) + in_zp)
  bool quant = false;
  float in_scale = 1.f;
  int in_zp = 0, in_qmin = 0, in_qmax = 255;
};

struct ModelError : public std::exception {
  int code;           // INFUR_E_MODEL_LOAD or INFUR_E_MODEL_INPUT_FORMAT
  std::string msg;
  ModelError(int c, std::string m) : code(c), msg(std::move(m)) {}
  const char* what() const noexcept override { return msg.c_str(); }
};

// Throws ModelError.
void lower_model(const OnnxGraph& g, LoweredModel& m);
// ResNet blocks with a projection shortcut compute relu(conv3(y) + downsample(x)): two 1x1 convolutions summed.
// This pass turns each such pair into ONE convolution with two sources (one accumulator, K = cin3 + cin_down),
// removing the shortcut tensor's round trip through memory.  Returns the number of pairs fused.
int fuse_projection_shortcuts(LoweredModel& m);
// Can a quantised model run as an int8 plan (u8 activation tensors in HBM, tcgen05.mma.kind::i8 convolutions; engine.h
// DeviceModel::i8)?  Every tensor must be u8, every convolution input except the RGB stem's must have zero point 0 (TMA's
// zero fill is then the padding value), weights must fit s8 once their zero point is subtracted, the pool must be 3x3 / s2 /
// p1, and a head convolution must not carry a residual.  `needed` (optional) masks the ops on the path of a computed head.
// On false, *why names the first obstacle; such models run with the integers carried exactly in fp16 instead.
bool int8_plan_eligible(const LoweredModel& m, const std::vector<char>* needed, std::string* why);
std::string describe(const LoweredModel& m);

}  // namespace infur
