// GTensor.hpp -- host-side mirror of the reference's tensor / quantiser surface for the inference hot path:
//   GTensor  (src/Tensor/GTensor.hpp:168-490): shape, type, ONE device blob data||gama, hQuant, GetDataX(), gama_T()
//   QUANT_CARD (src/CLI_params.hpp:509-554) + Init4Neuron (src/Tensor/GeQuant.cpp:1186-1285): quant-type selection by
//              neuron-name substring from the JSON "quantizer" block
//   GeQuant  (src/Tensor/GeQuant.hpp:94,115; GeQuant.cpp:23-81, 107-137, 830-905): MakeInstance, LowBit_worker
// Host code only orchestrates: every byte of arithmetic happens in the CUDA kernels behind include/kf_device.h.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "../Utils/json_lite.hpp"
#include "kf_device.h"

namespace koifish {

// g_float.hpp:84-117 (the subset on the hot path)
enum class typNUMBER : uint8_t {
    BF16, F8E5M2, Q4, Q2, T_SIGN, T_BINARY,
    Q4_NF,  /* Q4 under QUANT_MODE::RTNf: NormalFloat4 + per-row LUT */
    Q4_AWQ  /* Q4 under QUANT_MODE::AWQ: the vendor layout, stored [in_features][out_features] (GeQuant::ExTensor, GeQuant.cpp:144-200) */
};
int kfType(typNUMBER t);               // -> KF_T_*
const char* typName(typNUMBER t);
double BitPE(typNUMBER t);             // bits per element (src/Utils/GST_float.cpp:51)

enum QUANT_YYANG_ { I_OFF, I_01, I_11, I_TERNARY };           // src/CLI_params.hpp:500-505
enum QUANT_MODE { NO_QUANT, RTN, RTNf, F8Ex, AWQ };           // the modes Init4Neuron can select (GeQuant.cpp:1271-1281)

struct QUANT_CARD {
    int default_bits  = 4;
    int T_group       = 128;
    float T_errQ      = 0.3f;
    bool isSymmetric  = false;
    bool isZeroPoint  = false;
    bool has_filterQ  = false;  // "filterQ" present in the matched entry (parsed by the reference, unused by its isPass)
    bool isVendorQuant = false;
    QUANT_YYANG_ yyang = I_OFF;
    QUANT_MODE type    = NO_QUANT;
    std::string matched_key;

    // name: neuron / tensor name, e.g. "model.layers.3.self_attn.q_proj.weight"; jQuant: the JSON "quantizer" object.
    // Keys are matched as substrings of the name; '#'-prefixed keys are comments.  Returns type != NO_QUANT.
    bool Init4Neuron(const std::string& name, const JSON& jQuant);
    bool isPass() const { return type == NO_QUANT; }
    // storage type after quantisation: bit2typ(), GeQuant.cpp:127-137 + Bits2Type g_float.hpp:177-192
    typNUMBER tpQuant() const;
    int kfMode() const;  // KF_Q_*
};

struct Fish;
class GeQuant;
using hQUANT = std::shared_ptr<GeQuant>;

class GTensor : public std::enable_shared_from_this<GTensor> {
   public:
    enum GAMA_TYPE { GAMA, R_SCALE, C_SCALE, ZERO, STEP };  // src/Tensor/GTensor.hpp (gama_T selector)
    std::string name;
    int ne[2] = {0, 0};  // ne[0] = rows (N_out), ne[1] = cols (K_in)
    typNUMBER type = typNUMBER::BF16;
    void* data     = nullptr;  // device blob: data || gama  (GTensor.cpp:1017)
    size_t szData = 0, szGama = 0;
    int qBias = 0;
    hQUANT hQuant;
    kf_ctx* ctx = nullptr;

    GTensor(kf_ctx* ctx, const std::string& name, int rows, int cols);
    ~GTensor();
    size_t size() const { return (size_t)ne[0] * ne[1]; }
    size_t nByte() const { return szData + szGama; }
    // bf16 array directly after the packed bytes: [R_SCALE ne0][C_SCALE ne1][ZERO nG][STEP nG]  (GTensor.cpp:456-510)
    uint16_t* gama_T(GAMA_TYPE t = GAMA) const;
    int nGroup() const;
    kf_tensor_desc Desc() const;
    // Allocate as plain bf16 and upload / fill
    int Alloc(typNUMBER tp, int group);
    // vendor AWQ (GeQuant::ExTensor, GeQuant.cpp:144-200: hBase -> .qweight, aux .scales / .qzeros): ONE blob
    //   qweight int32 [cols][rows / 8]  ||  qzeros int32 [cols / 128][rows / 8]  ||  scales fp16 [cols / 128][rows]      (rows = out, cols = in)
    // szData = the qweight bytes, szGama = qzeros + scales.
    int AllocAWQ();
    size_t awqZeroBytes() const { return (size_t)(ne[1] / 128) * (ne[0] / 8) * 4; }
    int SetBF16FromDevice(const void* bf16_dev);  // un-quantised tensors (norms, bf16 embed)
    // GTensor::GetDataX (quantizer.cu:249-392): dequantise to a caller-provided bf16 device buffer (test hook)
    int GetDataX(void* out_bf16_dev) const;
};
using hGTensor = std::shared_ptr<GTensor>;

class GeQuant {
   public:
    QUANT_CARD params;
    int bits = 4, qMin = 0, qMax = 15, qBias = 0;
    explicit GeQuant(const QUANT_CARD& card);  // code ranges: GeQuant.cpp:107-124
    // MakeInstance (GeQuant.cpp:23-81): nullptr when the card does not quantise this neuron
    static hQUANT MakeInstance(const std::string& neuron_name, const JSON& jQuant);
    // LowBit_worker (GeQuant.cpp:830-905): quantise + pack `srcData` (bf16 [rows, cols]) into tensor->data / gama.
    // flag & 0x100: srcData is a device pointer (the only mode the product uses for big tensors); otherwise host.
    // Returns 0 on success, KF_ERR_* otherwise.
    int LowBit_worker(const hGTensor& tensor, const void* srcData, int flag);
};

}  // namespace koifish
