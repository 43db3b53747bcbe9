// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE (see oracle/README.md).
//
// A C-ABI door onto the UNMODIFIED reference slice runtime.  The reference source
// is #included where it lies (REF_TP_PATH = /root/reference/distllm/tensor_processor.cpp),
// so this is the same translation unit the reference's `llm` module is built from; the
// functions below only call what that file defines:
//   TransformerSlice            tensor_processor.cpp:1488-1562
//   get_inputs                  tensor_processor.cpp:1717-1784
//   get_llm_output              tensor_processor.cpp:1787-1892
//   sample_next_token           tensor_processor.cpp:1894-1908
//   llama_tokenize              tensor_processor.cpp:1700-1714
// What the door adds over the CPython module: raw float buffers instead of Python lists,
// and n_threads / n_ctx as arguments (the module hard-codes 3 / 512,
// tensor_processor.cpp:1997-2006; vendor examples/common.h:24-31).
#include REF_TP_PATH

extern "C" {

void * ref_slice_load(const char * path, int n_threads, int n_ctx) {
    gpt_params params;
    if (n_ctx > 0) params.n_ctx = n_ctx;
    TransformerSlice * s = new TransformerSlice(std::string(path), params, n_threads > 0 ? n_threads : 3);
    return (void *) s;
}

int ref_slice_n_embd(void * h) { return ((TransformerSlice *) h)->get_n_embd(); }

// in: [n_tokens][n_embd] f32, out: same shape.  Appends at the slice's internal n_past.
int ref_slice_forward(void * h, const float * in, int n_tokens, float * out) {
    TransformerSlice * s = (TransformerSlice *) h;
    const int n_embd = s->get_n_embd();
    std::vector<float> x(in, in + (size_t) n_tokens * n_embd), y;
    int rc = s->forward(x, y);
    if (rc != 0) return rc;
    if (y.size() != x.size()) return -2;
    memcpy(out, y.data(), y.size() * sizeof(float));
    return 0;
}

void ref_slice_clear(void * h) { ((TransformerSlice *) h)->clear_context(); }
void ref_slice_free(void * h)  { delete (TransformerSlice *) h; }

int ref_embed(const char * extra_path, const int * tokens, int n_tokens, float * out, int n_threads) {
    std::vector<llama_token> t(tokens, tokens + n_tokens);
    std::vector<float> e = get_inputs(std::string(extra_path), t.data(), n_tokens, n_threads > 0 ? n_threads : 3);
    memcpy(out, e.data(), e.size() * sizeof(float));
    return (int) e.size();
}

int ref_logits(const char * extra_path, const float * emb, int n_values, int all_logits, float * out) {
    std::vector<float> e(emb, emb + n_values);
    std::vector<float> l = get_llm_output(std::string(extra_path), e, all_logits != 0);
    memcpy(out, l.data(), l.size() * sizeof(float));
    return (int) l.size();
}

int ref_next_token(const char * extra_path, const float * emb, int n_values) {
    std::vector<float> e(emb, emb + n_values);
    return (int) sample_next_token(std::string(extra_path), e);
}

int ref_tokenize(const char * extra_path, const char * prompt, int * out, int cap) {
    llama_load_tensors_map tensors_map;
    my_file_loader loader(extra_path, tensors_map);
    std::vector<llama_token> t = llama_tokenize(loader.vocab, std::string(prompt), true);
    int n = (int) t.size() < cap ? (int) t.size() : cap;
    for (int i = 0; i < n; i++) out[i] = t[i];
    return (int) t.size();
}

}  // extern "C"
