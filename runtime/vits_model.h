// C++ drop-in for the reference runtime's `wetts::VitsModel` (runtime/core/model/vits_model.h:30-66) over the
// libwetts_b200 C ABI: same public methods and streaming semantics, no ONNXRuntime.  Where the reference passes
// `Ort::Value` tensors, this class passes plain float vectors in the same layout (z is time-major [L][192], the
// encoder graph's output layout, export_onnx.py:93-120).  Host-side code only: every tensor operation is a CUDA
// kernel behind include/wetts_b200.h; there is no CPU compute path.
#ifndef WETTS_B200_RUNTIME_VITS_MODEL_H_
#define WETTS_B200_RUNTIME_VITS_MODEL_H_

#include <cstdint>
#include <random>
#include <string>
#include <vector>

#include "wetts_b200.h"

namespace wetts {

const int kUpsampleRate = 256;   // vits_model.h:27

class VitsModel {
 public:
  // `weights_path`: flat weights file written by wetts_b200/flat.py (config + every checkpoint tensor).
  explicit VitsModel(const std::string& weights_path, int chunk_size = 40, int pad_size = 10, int device = 0,
                     uint64_t seed = 1234);
  ~VitsModel();
  VitsModel(const VitsModel&) = delete;
  VitsModel& operator=(const VitsModel&) = delete;

  // ids -> z (time-major [L][hidden_dim], masked), scales {0.667, 1.0, 0.8} as vits_model.cc:51
  std::vector<float> ForwardEncoder(const std::vector<int64_t>& phonemes, int sid);
  // z [L][hidden_dim] -> audio * 32767 (vits_model.cc:71-87)
  void ForwardDecoder(const std::vector<float>& z, int sid, std::vector<float>* audio);
  // non-stream call: ForwardEncoder then ForwardDecoder (vits_model.cc:89-93)
  void Forward(const std::vector<int64_t>& phonemes, int sid, std::vector<float>* audio);

  // stream call: encode once, then StreamDecode chunk by chunk until it returns true (vits_model.cc:127-153)
  void SetInput(const std::vector<int64_t>& phonemes, int sid);
  bool StreamDecode(std::vector<float>* audio);
  void SplitToChunks(const std::vector<float>& z);
  void Depadding(int chunk_id, int num_chunks, int chunk_size, int pad, int upsample, std::vector<float>* audio);

  int hidden_dim() const { return hidden_dim_; }
  int sampling_rate() const { return sampling_rate_; }
  int num_chunks() const { return static_cast<int>(z_chunks_.size()); }

 private:
  void EnsureWorkspace(size_t bytes);
  void Check(int status, const char* what) const;

  wetts_vits_t handle_ = nullptr;
  wetts_vits_config cfg_{};
  int sampling_rate_ = 22050;
  void* stream_ = nullptr;       // cudaStream_t
  void* workspace_ = nullptr;    // device, grow-only
  size_t workspace_bytes_ = 0;
  std::mt19937_64 rng_;

  int hidden_dim_ = 192;
  int chunk_size_ = 40;                       // stream decoder chunk size
  int pad_size_ = 10;                         // stream decoder pad size
  int sid_ = 0;                               // stream input sid
  int cur_ = 0;                               // stream synthesis index
  std::vector<std::vector<float>> z_chunks_;  // stream decoder z chunks
};

// Reads the flat weights file; returns false and fills *error on failure.
struct FlatTensor {
  std::string name;
  std::vector<int64_t> dims;
  std::vector<float> data;
};
bool LoadFlatWeights(const std::string& path, wetts_vits_config* cfg, int* sampling_rate,
                     std::vector<FlatTensor>* tensors, std::string* error);

}  // namespace wetts

#endif  // WETTS_B200_RUNTIME_VITS_MODEL_H_
