// GTensor.cpp -- host-side tensor + quantiser bookkeeping (see GTensor.hpp for the reference interfaces mirrored here).
#include "GTensor.hpp"

#include <algorithm>
#include <cctype>
#include <cstring>
#include <stdexcept>

namespace koifish {

int kfType(typNUMBER t) {
    switch (t) {
        case typNUMBER::BF16: return KF_T_BF16;
        case typNUMBER::F8E5M2: return KF_T_F8E5M2;
        case typNUMBER::Q4: return KF_T_Q4;
        case typNUMBER::Q2: return KF_T_Q2;
        case typNUMBER::T_SIGN: return KF_T_SIGN;
        case typNUMBER::T_BINARY: return KF_T_BINARY;
        case typNUMBER::Q4_NF: return KF_T_NF4;
        case typNUMBER::Q4_AWQ: return KF_T_AWQ4;
    }
    return KF_T_BF16;
}
const char* typName(typNUMBER t) {
    switch (t) {
        case typNUMBER::BF16: return "BF16";
        case typNUMBER::F8E5M2: return "F8E5M2";
        case typNUMBER::Q4: return "Q4";
        case typNUMBER::Q2: return "Q2";
        case typNUMBER::T_SIGN: return "T_SIGN";
        case typNUMBER::T_BINARY: return "T_BINARY";
        case typNUMBER::Q4_NF: return "Q4(NF4)";
        case typNUMBER::Q4_AWQ: return "Q4(AWQ)";
    }
    return "?";
}
double BitPE(typNUMBER t) {
    switch (t) {
        case typNUMBER::BF16: return 16;
        case typNUMBER::F8E5M2: return 8;
        case typNUMBER::Q4:
        case typNUMBER::Q4_AWQ:
        case typNUMBER::Q4_NF: return 4;
        case typNUMBER::Q2:
        case typNUMBER::T_SIGN: return 2;
        case typNUMBER::T_BINARY: return 1;
    }
    return 16;
}

static std::string lower(std::string s) {
    std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return (char)std::tolower(c); });
    return s;
}
static bool has_ci(const std::string& hay, const char* needle) { return lower(hay).find(lower(needle)) != std::string::npos; }

// QUANT_CARD::Init4Neuron, reference src/Tensor/GeQuant.cpp:1186-1285
bool QUANT_CARD::Init4Neuron(const std::string& name, const JSON& jQuant) {
    type = NO_QUANT;
    if (!jQuant.is_object() || jQuant.empty()) return false;
    if (jQuant.contains("VendorQuant")) isVendorQuant = true;
    if (const JSON* g = jQuant.find("group_size")) T_group = g->as_int(T_group);  // default group size (:1207-1209)
    for (const auto& kv : jQuant.obj) {
        const std::string& k = kv.first;
        if (k.empty() || k[0] == '#' || k == "debug") continue;  // '#'-prefixed keys are comments (:1226)
        if (name.find(k) == std::string::npos) continue;          // G_Has_(name, {k})
        const JSON& jQ = kv.second;
        if (!jQ.is_object()) continue;
        matched_key      = k;
        std::string info = jQ.contains("quant_method") ? jQ.at("quant_method").as_string() : "";
        if (const JSON* b = jQ.find("bits")) default_bits = b->as_int(default_bits);
        if (info == "bitnet") {
            // BLOCK_at_MATRIX ternary (:1240-1247): one group = whole matrix.  Not on the B200 path.
            throw std::runtime_error("quant_method 'bitnet' (BLOCK_at_MATRIX) is outside the B200 hot path scope");
        } else if (info == "yyang") {
            yyang  = default_bits == 1 ? I_01 : I_TERNARY;
            T_errQ = 1.0f;
        } else {
            if (default_bits != 4 && default_bits != 1 && default_bits != 2 && default_bits != 8) default_bits = 4;
            T_errQ = default_bits == 4 ? 0.3f : default_bits == 3 ? 0.4f : 0.7f;
        }
        if (const JSON* g = jQ.find("group_size")) T_group = g->as_int(T_group);
        if (T_group <= 0 || T_group >= 102400) throw std::runtime_error("quantizer: bad group_size");
        if (const JSON* z = jQ.find("zero_point")) isZeroPoint = z->as_bool(isZeroPoint);
        // "filterQ" is read into the card by the reference too (GeQuant.cpp:1231-1233) and then never consulted: QUANT_CARD::isPass
        // (:1286-1296) returns `type == NO_QUANT`, the name filter below it is commented out.  Same here: accepted, without effect.
        has_filterQ = jQ.find("filterQ") != nullptr;
        if (has_ci(info, "AWQ"))
            type = AWQ;
        else if (has_ci(info, "RTN") || has_ci(info, "bitnet"))
            type = RTN;
        else if (has_ci(info, "yyang"))
            type = RTN;
        else
            type = default_bits == 8 ? F8Ex : RTNf;
    }
    return type != NO_QUANT;
}
typNUMBER QUANT_CARD::tpQuant() const {
    if (type == F8Ex) return typNUMBER::F8E5M2;
    if (type == RTNf) return typNUMBER::Q4_NF;
    if (type == AWQ) return typNUMBER::Q4_AWQ;
    if (yyang == I_TERNARY) return typNUMBER::T_SIGN;  // bit2typ(), GeQuant.cpp:127-137
    switch (default_bits) {
        case 4: return typNUMBER::Q4;
        case 2: return typNUMBER::Q2;
        case 1: return typNUMBER::T_BINARY;
    }
    return typNUMBER::BF16;
}
int QUANT_CARD::kfMode() const {
    if (yyang != I_OFF) return KF_Q_YYANG;
    return isSymmetric ? KF_Q_RTN_SYM : KF_Q_RTN_ASYM;
}

// ------------------------------------------------------------------------------------------------ GTensor
GTensor::GTensor(kf_ctx* c, const std::string& n, int rows, int cols) : name(n), ctx(c) { ne[0] = rows, ne[1] = cols; }
GTensor::~GTensor() {
    if (data && ctx) kf_free(ctx, data);
}
int GTensor::nGroup() const {
    if (!hQuant || szGama == 0) return 0;
    return (int)(size() / hQuant->params.T_group);
}
uint16_t* GTensor::gama_T(GAMA_TYPE t) const {
    if (!data || szGama == 0) return nullptr;
    uint16_t* g0 = reinterpret_cast<uint16_t*>((uint8_t*)data + szData);
    switch (t) {
        case GAMA:
        case R_SCALE: return g0;
        case C_SCALE: return g0 + ne[0];
        case ZERO: return g0 + ne[0] + ne[1];
        case STEP: return g0 + ne[0] + ne[1] + nGroup();
    }
    return g0;
}
kf_tensor_desc GTensor::Desc() const {
    kf_tensor_desc d;
    d.data_dev = data;
    d.gama_dev = szGama ? (const void*)((const uint8_t*)data + szData) : nullptr;
    d.rows = ne[0], d.cols = ne[1];
    d.type  = kfType(type);
    d.group = hQuant ? hQuant->params.T_group : 128;
    d.qbias = qBias;
    d.zero_dev = nullptr, d.step_dev = nullptr;
    if (type == typNUMBER::Q4_AWQ) {  // include/kf_device.h KF_T_AWQ4: data_dev = qweight, zero_dev = qzeros, step_dev = scales
        d.gama_dev = nullptr, d.group = 128;
        if (data) {
            d.zero_dev = (const uint8_t*)data + szData;
            d.step_dev = (const uint8_t*)data + szData + awqZeroBytes();
        }
    }
    return d;
}
int GTensor::AllocAWQ() {
    if (data) {
        kf_free(ctx, data);
        data = nullptr;
    }
    // whole 128-row groups, whole 32-column blocks: every section of the blob stays 16-byte aligned (awq.cu reads 8 scales as one uint4)
    if (ne[1] % 128 || ne[0] % 32) return KF_ERR_BAD_ARG;
    type   = typNUMBER::Q4_AWQ;
    szData = (size_t)ne[1] * (ne[0] / 8) * 4;
    szGama = awqZeroBytes() + (size_t)(ne[1] / 128) * ne[0] * 2;
    return kf_malloc(ctx, szData + szGama + 16, &data);
}
int GTensor::Alloc(typNUMBER tp, int group) {
    if (tp == typNUMBER::Q4_AWQ) return AllocAWQ();
    if (data) {
        kf_free(ctx, data);
        data = nullptr;
    }
    type   = tp;
    szData = kf_quant_data_bytes(ne[0], ne[1], kfType(tp));
    szGama = kf_quant_gama_bytes(ne[0], ne[1], kfType(tp), group);
    // keep gama 16-byte aligned behind the packed bytes (szData is a multiple of 16 for every supported shape)
    return kf_malloc(ctx, szData + szGama + 16, &data);
}
int GTensor::SetBF16FromDevice(const void* src) {
    int rc = Alloc(typNUMBER::BF16, 0);
    if (rc) return rc;
    return kf_d2d(ctx, data, src, szData);
}
int GTensor::GetDataX(void* out) const {
    kf_tensor_desc d = Desc();
    return kf_dequant(ctx, &d, out);
}

// ------------------------------------------------------------------------------------------------ GeQuant
GeQuant::GeQuant(const QUANT_CARD& card) : params(card) {
    bits = card.default_bits;
    if (params.yyang != I_OFF) {  // GeQuant.cpp:107-117
        if (bits == 2) {
            qMax = 1, qMin = -1, qBias = 1;
            params.isSymmetric = true;
        } else {
            qMax = 1, qMin = 0, qBias = 0;
            params.isSymmetric = false;
        }
    } else if (params.isSymmetric) {
        qMin = -(1 << (bits - 1)), qMax = (1 << (bits - 1)) - 1, qBias = -qMin;
    } else {
        qMin = 0, qMax = (1 << bits) - 1, qBias = 0;
    }
}
hQUANT GeQuant::MakeInstance(const std::string& neuron_name, const JSON& jQuant) {
    QUANT_CARD card;
    if (!card.Init4Neuron(neuron_name, jQuant)) return nullptr;
    switch (card.type) {
        case RTN:
        case F8Ex: break;
        case RTNf:
            if (card.default_bits != 4)  // RT_NormalF also accepts 3 bits (NF3), which no PackedQ storage type carries
                throw std::runtime_error("quantizer entry '" + card.matched_key + "': NormalFloat (no quant_method) is built for bits = 4 only");
            break;
        case AWQ:  // Q_AWQ (GeQuant.cpp:989-1013): 4-bit codes, 128-row groups, explicit .qzeros / .scales tensors from the vendor checkpoint
            if (card.default_bits != 4 || card.T_group != 128)
                throw std::runtime_error("quantizer entry '" + card.matched_key + "': the vendor AWQ layout is built for bits = 4, group_size = 128");
            break;
        default: return nullptr;
    }
    if (card.type == RTN && card.default_bits == 1 && card.yyang == I_OFF)
        throw std::runtime_error("1-bit weights need quant_method 'yyang' (GeQuant::Core, GeQuant.cpp:909)");
    if (card.type == RTN && card.default_bits == 8) throw std::runtime_error("8-bit RTN is not a reference mode; omit quant_method for F8Ex");
    return std::make_shared<GeQuant>(card);
}
int GeQuant::LowBit_worker(const hGTensor& t, const void* srcData, int flag) {
    if (!t || !srcData) return KF_ERR_BAD_ARG;
    const typNUMBER tp = params.tpQuant();
    const int rows = t->ne[0], cols = t->ne[1];
    // AWQ tensors arrive packed from vendor checkpoints (Fish::SetTensorAWQ); like the reference there is no AWQ quantiser: kf_quantize
    // refuses the type with a message
    if (tp == typNUMBER::Q4_AWQ) return kf_quantize(t->ctx, srcData, rows, cols, KF_T_AWQ4, 128, KF_Q_RTN_ASYM, const_cast<void*>(srcData) /* never written */, nullptr, nullptr);
    int rc = t->Alloc(tp, params.T_group);
    if (rc) return rc;
    void* src_dev = const_cast<void*>(srcData);
    void* staged  = nullptr;
    if (!(flag & 0x100)) {  // host source: upload, then quantise on the device (there is no CPU quantiser in the product)
        rc = kf_malloc(t->ctx, t->size() * 2, &staged);
        if (rc) return rc;
        rc = kf_h2d(t->ctx, staged, srcData, t->size() * 2);
        if (rc) {
            kf_free(t->ctx, staged);
            return rc;
        }
        src_dev = staged;
    }
    void* gama = t->szGama ? (void*)((uint8_t*)t->data + t->szData) : nullptr;
    int qb     = 0;
    rc         = kf_quantize(t->ctx, src_dev, rows, cols, kfType(tp), params.T_group, params.kfMode(), t->data, gama, &qb);
    t->qBias   = qb;
    if (staged) {
        kf_ctx_sync(t->ctx);
        kf_free(t->ctx, staged);
    }
    return rc;
}

}  // namespace koifish
