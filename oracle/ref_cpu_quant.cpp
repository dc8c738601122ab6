// ref_cpu_quant.cpp -- TEST INFRASTRUCTURE ONLY.  Runs the reference's own host code on caller buffers: the CPU packers (below), and, at the end of
// the file, QUANT_CARD::Vendor2JSONx / Init4Neuron, CHAT_SAMPLER::toChatML / InitPrefillTemplate and the sampler (LogitsInfo, src/Manifold/GoPT.cpp).
// The CPU packers -- GeQuant::RTN_x (4- / 2-bit, asymmetric, symmetric, ternary
// yyang) and GeQuant::YinYang (1-bit), reference src/Tensor/GeQuant.cpp:428-533, 536-628 -- on caller buffers, so that the oracle's restatement
// (kfo_quantize) can be pinned to the code the reference compiles.  oracle/Makefile builds the reference's GeQuant.cpp and GTensor.cpp where they
// lie into objects and links them with this shim into oracle/_ref/libkoifish_refcpu.so; every symbol of the rest of the framework those two files
// mention (Fish, CUDA runtime, optimizers ...) is bound to 0 at link time and never reached: the packers touch only the members set up below.
// The GeQuant object is built by the reference's own constructor (GeQuant.cpp:83-124: code ranges from the quantizer card), the GTensor by its
// default constructor; both are deliberately leaked (their destructors belong to the framework).  Nothing of the reference is copied.
#include <cstring>

#include "Manifold/Fish.hpp"
#include "Manifold/GoPT.hpp"
#include "Manifold/Neuron.hpp"
#include "Tensor/GeQuant.hpp"
#include "Tensor/GTensor.hpp"
#include "Utils/GST_util.hpp"

double SUM::tQuant = 0, SUM::tF8Ex = 0, SUM::tLowBit = 0;  // defined in src/Utils/GST_util.cpp, which drags the framework in

namespace {
struct QuantShim : public GeQuant {  // reach the protected working buffers of the packers
    // the reference's own constructor (GeQuant.cpp:83-124) derives the code range qMin / qMax / qBias from the card
    QuantShim(QUANT_CARD& card) : GeQuant("refcpu", nullptr, card, 0x0) {}
    void Target(hBITARR data, floatGama* gama_) {
        // the constructor sized its own buffers for the framework's tensors (max(nGroup, T_group) * 2^bits + ...), which a short-and-wide test matrix
        // can exceed: the packers write into the caller's buffers instead (the constructor's allocations are leaked with the object)
        isGPU      = false;
        quant_data = data;
        gama       = gama_;
    }
};
struct TensorShim : public GTensor {  // hQuant is protected: GTensor::gama_T asks it for the group count
    void SetHost(void* p) { host_data = p; }
    void Setup(int rows, int cols, GeQuant* q) {
        for (int i = 0; i < N_DIMS; i++) ne[i] = 1;
        ne[0] = rows, ne[1] = cols;
        shape  = {rows, cols};
        hQuant = std::shared_ptr<GeQuant>(q, [](GeQuant*) {});
    }
};
}  // namespace

// mode: 0 asymmetric RTN, 1 symmetric RTN, 2 yyang (2-bit ternary through RTN_x, 1-bit through YinYang), 3 NormalFloat4 (RTN_x -> RT_NormalF ->
// _row_lut, GeQuant.cpp:706-752: per-row codebook + MSB-first bitstream).  w: bf16 [rows][cols]; data_out: rows * cols * bits / 8 bytes;
// gama_out: rows + cols + 2 * (rows * cols / group) bf16 ([R_SCALE][C_SCALE][ZERO][STEP]), or rows + cols + 16 * rows for mode 3 ([..][..][LUT]).
extern "C" int refcpu_quantize(const void* w_bf16, int rows, int cols, int bits, int group, int mode, void* data_out, void* gama_out, int* qbias_out) {
    if (!w_bf16 || !data_out || !gama_out || rows <= 0 || cols <= 0 || group <= 0 || ((size_t)rows * cols) % group) return -1;
    if (!(bits == 4 || bits == 2 || bits == 1) || (bits == 1 && mode != 2) || (mode == 3 && bits != 4)) return -2;
    QUANT_CARD card;
    card.default_bits = bits, card.T_group = group, card.blockAt = BLOCK_at_GROUP, card.type = QUANT_MODE::RTN;
    card.isNormalFloat = mode == 3, card.norm = NORMAL_MODE::NO_NORMAL;  // NO_NORMAL is forced at GeQuant.cpp:844-852
    card.isSymmetric   = mode == 1;
    card.yyang         = mode == 2 ? (bits == 1 ? QUANT_YYANG_::I_01 : QUANT_YYANG_::I_TERNARY) : QUANT_YYANG_::I_OFF;  // Init4Neuron, GeQuant.cpp:1248-1250
    card.spMost        = {rows, cols};
    auto* q = new QuantShim(card);
    q->Target((hBITARR)data_out, (floatGama*)gama_out);
    auto* t = new TensorShim();
    t->Setup(rows, cols, q);
    std::shared_ptr<GTensor> ht(t, [](GTensor*) {});
    if (bits == 1)
        q->YinYang(ht, w_bf16, 0);
    else
        q->RTN_x(ht, w_bf16, 0);
    if (qbias_out) *qbias_out = q->qBias;
    return 0;
}

// QUANT_CARD::Vendor2JSONx (reference src/Utils/CLI_params.cpp:240-262): an HF checkpoint's "quantization_config" block -> the "quantizer" block the
// reference builds from it.  JSON text in, JSON text out (the reference's own nlohmann dump); returns the length, or -1 when `cap` is too small.
extern "C" int refcpu_vendor2jsonx(const char* vendor_json, char* out, int cap) {
    if (!vendor_json || !out) return -1;
    const JSON jx     = JSON::parse(vendor_json);
    const std::string s = QUANT_CARD::Vendor2JSONx(jx).dump();
    if ((int)s.size() + 1 > cap) return -1;
    memcpy(out, s.c_str(), s.size() + 1);
    return (int)s.size();
}

// CHAT_SAMPLER::toChatML (reference src/Utils/CLI_params.cpp:2010-2031) over n (role, content) lines; text out as refcpu_vendor2jsonx
extern "C" int refcpu_tochatml(const char* const* roles, const char* const* contents, int n, int enable_thinking, char* out, int cap) {
    if (n < 0 || !out) return -1;
    CHAT_SAMPLER cs;
    cs.enable_thinking = enable_thinking != 0;
    std::vector<ChatML_samp> lines;
    for (int i = 0; i < n; i++) lines.emplace_back(std::string(roles[i]), std::string(contents[i]));
    const std::string s = cs.toChatML(lines);
    if ((int)s.size() + 1 > cap) return -1;
    memcpy(out, s.c_str(), s.size() + 1);
    return (int)s.size();
}
// CHAT_SAMPLER::InitPrefillTemplate (:1990-2008): the two printf templates (user only; system + user) for enable_thinking on / off, joined by '\x01'
extern "C" int refcpu_prefill_templates(int enable_thinking, char* out, int cap) {
    if (!out) return -1;
    auto* cfg = new CLI_params();  // leaked: its destructor belongs to the framework
    cfg->model.enable_thinking = enable_thinking != 0;
    CHAT_SAMPLER cs;
    cs.InitPrefillTemplate(cfg);
    const std::string s = cs.prompt_template + "\x01" + cs.system_prompt_template;
    if ((int)s.size() + 1 > cap) return -1;
    memcpy(out, s.c_str(), s.size() + 1);
    return (int)s.size();
}

// QUANT_CARD::Init4Neuron (reference src/Tensor/GeQuant.cpp:1186-1285): which quantiser the "quantizer" block selects for a tensor name.  The
// function asks its neuron for hFish->isAtPhase(P_CHAT_1) and copies hFish->config.distill; both objects are zero-filled storage of the right size
// (never constructed, never destroyed), and Fish::isAtPhase -- a one-line accessor of src/Manifold/Fish.cpp, which is not linked -- is answered here.
bool Fish::isAtPhase(LIFE_PHASE) const { return true; }  // the chat phase: the only effect is QUANT_CARD::isDequant4Generate
namespace {
struct NeuronAccess : public GeNeuron {
    static void SetFish(GeNeuron* n, Fish* f) { static_cast<NeuronAccess*>(n)->hFish = f; }
};
}  // namespace
// out[8] = {selected (0 / 1), type (QUANT_MODE), default_bits, T_group, yyang, isSymmetric, isZeroPoint, isVendorQuant}; *errq_out = T_errQ
extern "C" int refcpu_init4neuron(const char* tensor_name, const char* quantizer_json, int* out, float* errq_out) {
    if (!tensor_name || !quantizer_json || !out) return -1;
    static void* fish   = calloc(1, sizeof(Fish));
    static void* neuron = calloc(1, sizeof(GeNeuron));
    NeuronAccess::SetFish((GeNeuron*)neuron, (Fish*)fish);
    const JSON jq = JSON::parse(quantizer_json);
    QUANT_CARD card;
    const bool sel = card.Init4Neuron(tensor_name, jq, neuron, 0x0);
    out[0] = sel, out[1] = (int)card.type, out[2] = card.default_bits, out[3] = card.T_group, out[4] = (int)card.yyang, out[5] = card.isSymmetric,
    out[6] = card.isZeroPoint, out[7] = card.isVendorQuant;
    if (errq_out) *errq_out = card.T_errQ;
    return 0;
}

// The sampler of the chat loop: GeneratOnPrompt::Sample (reference src/Manifold/GoPT.cpp:614-630) = LogitsInfo::TopK (:632-640, TOPK_heap::Select
// :667-700) -> UpdateLogits (:751-766) -> TopP (:729-748) -> Qu_FlipCoin (:768-786, xorshift64* :594-600), on bf16 logits in host memory.
// LogitsInfo's constructor reads hFish->config.common.seed and the virtual hFish->nClass(): the Fish is zero-filled storage whose vtable pointer is
// aimed at a table where every slot answers with the vocabulary size (Itanium ABI: a slot is a plain function taking `this`).
namespace {
size_t g_vocab = 0;
size_t AnswerVocab(const void*) { return g_vocab; }
}  // namespace
extern "C" int refcpu_sample(const void* logits_bf16, int vocab, float temperature, int top_k, float top_p, unsigned long long* rng_state, int* n_pick_out) {
    if (!logits_bf16 || vocab < 4 || !rng_state || temperature <= 0.f || top_k < 2) return -1;
    static void* slots[2048];
    static void* fish = nullptr;
    if (!fish) {
        for (auto& s : slots) s = (void*)&AnswerVocab;
        fish           = calloc(1, sizeof(Fish));
        *(void***)fish = slots;
    }
    g_vocab = (size_t)vocab;
    auto* t = new TensorShim();
    t->SetHost(const_cast<void*>(logits_bf16));
    std::shared_ptr<GTensor> ht(t, [](GTensor*) {});
    LogitsInfo li(0, (const Fish*)fish, ht);
    li.rng_state = *rng_state;
    CHAT_SAMPLER samp;
    samp.temperature = temperature, samp.top_k = top_k, samp.top_p = top_p;
    const int nCanTopK = top_k < vocab ? top_k : vocab;  // GeneratOnPrompt ctor, GoPT.cpp:389
    li.TopK(nCanTopK);                                   // Sample(), GoPT.cpp:621-625
    li.UpdateLogits(samp);
    li.TopP(samp.top_p, nCanTopK);
    li.Qu_FlipCoin();
    *rng_state = li.rng_state;
    if (n_pick_out) *n_pick_out = li.nPick;
    return (int)li.qu;
}
