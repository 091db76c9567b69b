// psa_capi.cpp -- the reference-facing surface: gpu_run_program drop-in, input.txt / output.txt,
// and psa_run_files (what initiate_program does around the hot path, cpu_funcs.c:25-121).
#include "psa_host.h"
#include "psa_reference_abi.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cctype>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

static_assert(offsetof(ProgramData, weights) == 8, "ProgramData layout (program_data.h:6-11)");
static_assert(offsetof(ProgramData, seq1) == 40, "ProgramData layout (program_data.h:6-11)");
static_assert(offsetof(ProgramData, seq2) == 40 + PSA_SEQ1_CAPACITY + 1, "ProgramData layout");
static_assert(sizeof(Mutant) == 12 && offsetof(Mutant, ch) == 8, "Mutant layout (mutant.h:6-10)");
static_assert(sizeof(psa_mutant) == sizeof(Mutant), "psa_mutant mirrors Mutant");

namespace {

// gpu_run_program has no context argument: keep one per CUDA device, created on first use and
// kept for the life of the process (the reference allocates and frees per call instead).
std::mutex g_mu;
std::map<int, psa_context*> g_ctx;

[[noreturn]] void die(const char* what, const char* detail)
{
    // the reference's convention for every CUDA failure (cuda_funcs.cu:44-48)
    std::fprintf(stderr, "%s - %s\n", what, detail);
    std::exit(EXIT_FAILURE);
}

psa_context* context_for_current_device()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) die("Failed to query the CUDA device", cudaGetErrorString(cudaGetLastError()));
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_ctx.find(dev);
    if (it != g_ctx.end()) return it->second;
    psa_context* ctx = nullptr;
    int rc = psa_create(&ctx, &dev, 1);
    if (rc) die("Failed to create the GPU context", psa_strerror(rc));
    g_ctx[dev] = ctx;
    return ctx;
}

} // namespace

// C++ linkage on purpose: this is the symbol cpu_funcs.c:180 resolves.
double gpu_run_program(ProgramData* cpu_data, Mutant* returned_mutant, int first_offset, int last_offset)
{
    if (!cpu_data || !returned_mutant) die("gpu_run_program", "null argument");
    const double none = cpu_data->is_max ? -INFINITY : INFINITY;
    returned_mutant->offset = -1;
    returned_mutant->char_offset = -1;
    returned_mutant->ch = '\0';
    if (last_offset <= first_offset) return none;         // the reference's caller never does this (cpu_funcs.c:178)
    psa_context* ctx = context_for_current_device();
    psa_result r;
    int rc = psa_search_range(ctx, cpu_data->weights, cpu_data->is_max, cpu_data->seq1,
                              (int64_t)std::strlen(cpu_data->seq1), cpu_data->seq2,
                              (int64_t)std::strlen(cpu_data->seq2), first_offset, last_offset, &r);
    if (rc) die("Failed to run the mutant-offset search on the GPU", psa_last_error(ctx));
    returned_mutant->offset = r.mutant.offset;
    returned_mutant->char_offset = r.mutant.char_offset;
    returned_mutant->ch = r.mutant.ch;
    return r.mutant.ch == '\0' ? none : r.score;           // cuda_funcs.cu:143-145
}

// -------------------------------------------------------------------------------------------------
// The six host primitives cpu_funcs.c imports from the reference's cuda_funcs.cu (cuda_funcs.h:44-61; nm -u of its
// object file: get_weight, strlen_gpu, is_swapable, get_pair_sign, get_substitute, get_hashtable_sign).  Exported
// under the same C++-mangled names so that the reference links with NO cuda_funcs.o at all (include/cuda_funcs.h,
// INTEGRATION.md 1b).  Host only; every answer comes from the resolver in psa_table.cpp, none from the reference's
// hashtable_cpu -- so its racy fill_hash (cpu_funcs.c:306-307) cannot reach a score any more.
// -------------------------------------------------------------------------------------------------
char get_hashtable_sign(char c1, char c2)
{
    // cuda_funcs.cu:424-439: the gap symbol pairs only with itself, anything else outside A..Z has no sign
    if (c1 == '-' || c2 == '-') return c1 == c2 ? '*' : '_';
    const int a = psa::symbol_index(c1), b = psa::symbol_index(c2);
    if (a < 0 || b < 0) return '\0';
    return psa::sign_of(a, b);
}

char get_pair_sign(char a, char b)
{
    // cuda_funcs.cu:495-502: group membership only; characters in no group (including '-') are '_' unless identical
    if (a == b) return '*';
    const int x = psa::symbol_index(a), y = psa::symbol_index(b);
    if (x < 0 || y < 0 || x == psa::kGap || y == psa::kGap) return '_';
    return psa::sign_of(x, y);
}

double get_weight(char sign, double* w)
{
    // cuda_funcs.cu:442-452
    return sign == '*' ? w[0] : sign == ':' ? -w[1] : sign == '.' ? -w[2] : sign == '_' ? -w[3] : 0.0;
}

char get_substitute(char c1, char c2, double* w, int is_max)
{
    // cuda_funcs.cu:310-421 as one lookup; undefined in the reference for symbols outside [A-Z-]: '\0' here
    const int a = psa::symbol_index(c1), b = psa::symbol_index(c2);
    if (a < 0 || b < 0 || !w) return '\0';
    const int s = psa::best_substitute(a, b, w, is_max != 0);
    return s < 0 ? '\0' : psa::symbol_char(s);
}

int is_swapable(Mutant* m1, Mutant* m2, double score1, double score2, int is_max)
{
    // cuda_funcs.cu:290-307: strictly better score, else on equal scores the lower (offset, char_offset)
    if (is_max ? score2 > score1 : score2 < score1) return 1;
    if (score2 != score1) return 0;
    if (m2->offset != m1->offset) return m2->offset < m1->offset;
    return m2->char_offset < m1->char_offset;
}

int strlen_gpu(char* str)
{
    int n = 0;
    while (str[n]) n++;
    return n;
}

extern "C" {

// C spellings of the same six, for ctypes / C callers
char   psa_get_hashtable_sign(char c1, char c2) { return get_hashtable_sign(c1, c2); }
char   psa_get_pair_sign(char a, char b) { return get_pair_sign(a, b); }
double psa_get_weight(char sign, const double* w) { return get_weight(sign, const_cast<double*>(w)); }
char   psa_get_substitute(char c1, char c2, const double* w, int is_max) { return get_substitute(c1, c2, const_cast<double*>(w), is_max); }
int    psa_is_swapable(const psa_mutant* m1, const psa_mutant* m2, double score1, double score2, int is_max)
{
    Mutant a = { m1->offset, m1->char_offset, m1->ch }, b = { m2->offset, m2->char_offset, m2->ch };
    return is_swapable(&a, &b, score1, score2, is_max);
}
int    psa_strlen(const char* str) { return strlen_gpu(const_cast<char*>(str)); }

double psa_gpu_run_program(void* program_data, void* returned_mutant, int first_offset, int last_offset)
{
    return gpu_run_program((ProgramData*)program_data, (Mutant*)returned_mutant, first_offset, last_offset);
}

// Whitespace-separated tokens, like the reference's fscanf("%lf %lf %lf %lf") + three fscanf("%s")
// (cpu_funcs.c:353-368), but without its fixed-size buffers.
int psa_read_input_file(const char* path, double weights[4], int* is_max, char** seq1, char** seq2)
{
    if (!path || !weights || !is_max || !seq1 || !seq2) return PSA_ERR_ARG;
    *seq1 = *seq2 = nullptr;
    FILE* f = std::fopen(path, "r");
    if (!f) return PSA_ERR_IO;
    int rc = PSA_ERR_IO;
    std::string tok[3];
    if (std::fscanf(f, "%lf %lf %lf %lf", &weights[0], &weights[1], &weights[2], &weights[3]) == 4) {
        int got = 0;
        for (; got < 3; got++) {
            int c;
            while ((c = std::fgetc(f)) != EOF && std::isspace(c)) {}
            if (c == EOF) break;
            do { tok[got].push_back((char)c); } while ((c = std::fgetc(f)) != EOF && !std::isspace(c));
        }
        if (got == 3) {
            *is_max = tok[2] == "maximum" ? 1 : 0;        // anything else is a minimum (cpu_funcs.c:365)
            *seq1 = (char*)std::malloc(tok[0].size() + 1);
            *seq2 = (char*)std::malloc(tok[1].size() + 1);
            if (*seq1 && *seq2) {
                std::memcpy(*seq1, tok[0].c_str(), tok[0].size() + 1);
                std::memcpy(*seq2, tok[1].c_str(), tok[1].size() + 1);
                rc = PSA_OK;
            } else {
                std::free(*seq1); std::free(*seq2);
                *seq1 = *seq2 = nullptr;
                rc = PSA_ERR_NOMEM;
            }
        }
    }
    std::fclose(f);
    return rc;
}

int psa_write_output_file(const char* path, const char* mutant, int offset, double score)
{
    if (!path || !mutant) return PSA_ERR_ARG;
    FILE* f = std::fopen(path, "w");
    if (!f) return PSA_ERR_IO;
    // "<mutant>\n<offset> <score>" with %g and no trailing newline (cpu_funcs.c:377)
    int ok = std::fprintf(f, "%s\n%d %g", mutant, offset, score) > 0;
    ok = (std::fclose(f) == 0) && ok;
    return ok ? PSA_OK : PSA_ERR_IO;
}

int psa_run_files(psa_context* ctx, const char* input_path, const char* output_path, psa_result* out)
{
    if (!ctx || !input_path || !output_path) return PSA_ERR_ARG;
    double w[4];
    int is_max = 0;
    char *seq1 = nullptr, *seq2 = nullptr;
    int rc = psa_read_input_file(input_path, w, &is_max, &seq1, &seq2);
    if (rc) return rc;
    const int64_t len1 = (int64_t)std::strlen(seq1), len2 = (int64_t)std::strlen(seq2);
    const int64_t q_off[2] = { 0, len2 };
    psa_result r;
    rc = psa_search_batch(ctx, w, is_max, seq1, len1, seq2, q_off, 1, &r);
    if (rc == PSA_OK) {
        // the mutant string: Seq2 with one character replaced (cpu_funcs.c:96-98); if no mutation
        // exists the reference indexes with -1 (undefined) -- here Seq2 is written unchanged
        std::string mut(seq2);
        if (r.mutant.char_offset >= 0) mut[(size_t)r.mutant.char_offset] = r.mutant.ch;
        rc = psa_write_output_file(output_path, mut.c_str(), r.mutant.offset, r.score);
        if (out) *out = r;
    }
    std::free(seq1); std::free(seq2);
    return rc;
}

void psa_free(void* p) { std::free(p); }

int psa_read_query_file(const char* path, char** seq2s, int64_t** q_off, int32_t* nq)
{
    if (!path || !seq2s || !q_off || !nq) return PSA_ERR_ARG;
    *seq2s = nullptr; *q_off = nullptr; *nq = 0;
    FILE* f = std::fopen(path, "r");
    if (!f) return PSA_ERR_IO;
    std::string text;
    char buf[1 << 16];
    size_t got;
    while ((got = std::fread(buf, 1, sizeof(buf), f)) > 0) text.append(buf, got);
    std::fclose(f);
    // FASTA if any line starts with '>': header lines end the record before them; otherwise every token is a query
    bool fasta = false;
    for (size_t i = 0; i < text.size(); i++)
        if (text[i] == '>' && (i == 0 || text[i - 1] == '\n')) { fasta = true; break; }
    std::string cat;
    std::vector<int64_t> offs(1, 0);
    auto close_record = [&]() { if ((int64_t)cat.size() > offs.back()) offs.push_back((int64_t)cat.size()); };
    size_t i = 0;
    while (i < text.size()) {
        size_t e = text.find('\n', i);
        if (e == std::string::npos) e = text.size();
        if (fasta && text[i] == '>') close_record();
        else if (fasta && text[i] == ';') {}                      // old-style FASTA comment line
        else
            for (size_t k = i; k < e; k++) {
                const unsigned char c = (unsigned char)text[k];
                if (std::isspace(c)) { if (!fasta) close_record(); }
                else cat.push_back((char)c);
            }
        if (!fasta) close_record();
        i = e + 1;
    }
    close_record();
    const size_t n = offs.size() - 1;
    if (n > 0x7FFFFFFFu) return PSA_ERR_ARG;
    *seq2s = (char*)std::malloc(cat.size() + 1);
    *q_off = (int64_t*)std::malloc(sizeof(int64_t) * (n + 1));
    if (!*seq2s || !*q_off) { std::free(*seq2s); std::free(*q_off); *seq2s = nullptr; *q_off = nullptr; return PSA_ERR_NOMEM; }
    std::memcpy(*seq2s, cat.data(), cat.size());
    (*seq2s)[cat.size()] = '\0';
    std::memcpy(*q_off, offs.data(), sizeof(int64_t) * (n + 1));
    *nq = (int32_t)n;
    return PSA_OK;
}

int psa_run_query_file(psa_context* ctx, const char* input_path, const char* queries_path, const char* output_path, int32_t* nq_out)
{
    if (!ctx || !input_path || !queries_path || !output_path) return PSA_ERR_ARG;
    if (nq_out) *nq_out = 0;
    double w[4];
    int is_max = 0;
    char *seq1 = nullptr, *seq2 = nullptr, *qs = nullptr;
    int64_t* q_off = nullptr;
    int32_t nq = 0;
    int rc = psa_read_input_file(input_path, w, &is_max, &seq1, &seq2);
    if (rc == PSA_OK) rc = psa_read_query_file(queries_path, &qs, &q_off, &nq);
    if (rc == PSA_OK && nq == 0) rc = PSA_ERR_IO;
    std::vector<psa_result> res((size_t)std::max(nq, 1));
    std::string mutants;
    if (rc == PSA_OK) {
        mutants.resize((size_t)q_off[nq]);
        rc = psa_search_batch_mutants(ctx, w, is_max, seq1, (int64_t)std::strlen(seq1), qs, q_off, nq, res.data(), &mutants[0]);
    }
    if (rc == PSA_OK) {
        std::string text;
        for (int32_t q = 0; q < nq; q++) {
            char tail[96];
            std::snprintf(tail, sizeof(tail), "\n%d %g", res[q].mutant.offset, res[q].score);     // cpu_funcs.c:377
            if (q) text += "\n";
            text.append(mutants, (size_t)q_off[q], (size_t)(q_off[q + 1] - q_off[q]));
            text += tail;
        }
        FILE* o = std::fopen(output_path, "w");
        if (!o) rc = PSA_ERR_IO;
        else {
            bool wrote = std::fwrite(text.data(), 1, text.size(), o) == text.size();
            wrote = (std::fclose(o) == 0) && wrote;
            if (!wrote) rc = PSA_ERR_IO;
        }
    }
    if (rc == PSA_OK && nq_out) *nq_out = nq;
    std::free(seq1); std::free(seq2); psa_free(qs); psa_free(q_off);
    return rc;
}

int psa_run_files_all(psa_context* ctx, const char* input_path, const char* output_path, int* nblocks)
{
    if (!ctx || !input_path || !output_path) return PSA_ERR_ARG;
    if (nblocks) *nblocks = 0;
    FILE* f = std::fopen(input_path, "r");
    if (!f) return PSA_ERR_IO;
    std::vector<std::string> tok;
    {
        std::string cur;
        int c;
        while ((c = std::fgetc(f)) != EOF) {
            if (std::isspace(c)) { if (!cur.empty()) { tok.push_back(cur); cur.clear(); } }
            else cur.push_back((char)c);
        }
        if (!cur.empty()) tok.push_back(cur);
    }
    std::fclose(f);
    if (tok.size() < 7) return PSA_ERR_IO;                      // not even one complete block
    // every complete block is one problem of a pipelined list (psa_search_many): the copies and launch of one block overlap
    // the kernel of another, and on a multi-GPU context the blocks spread over the GPUs
    const size_t nb = tok.size() / 7;
    std::vector<double> weights(4 * nb);
    std::vector<int64_t> q_off(2 * nb);
    std::vector<psa_result> res(nb);
    std::vector<psa_problem> problems(nb);
    for (size_t b = 0; b < nb; b++) {
        const size_t k = 7 * b;
        bool ok = true;
        for (int i = 0; i < 4 && ok; i++) {
            char* end = nullptr;
            weights[4 * b + i] = std::strtod(tok[k + i].c_str(), &end);
            ok = end && *end == '\0' && end != tok[k + i].c_str();
        }
        if (!ok) return PSA_ERR_IO;
        const std::string &s1 = tok[k + 4], &s2 = tok[k + 5];
        q_off[2 * b] = 0; q_off[2 * b + 1] = (int64_t)s2.size();
        psa_problem& p = problems[b];
        p.weights = &weights[4 * b];
        p.is_max = tok[k + 6] == "maximum" ? 1 : 0;
        p.nq = 1;
        p.seq1 = s1.c_str(); p.len1 = (int64_t)s1.size();
        p.seq2s = s2.c_str(); p.q_off = &q_off[2 * b];
        p.out = &res[b];
        p.status = PSA_OK; p.reserved = 0;
    }
    if (int rc = psa_search_many(ctx, problems.data(), (int32_t)nb, 0)) return rc;
    std::string text;
    int done = 0;
    for (size_t b = 0; b < nb; b++) {
        const psa_result& r = res[b];
        std::string mut(tok[7 * b + 5]);
        if (r.mutant.char_offset >= 0) mut[(size_t)r.mutant.char_offset] = r.mutant.ch;
        char tail[96];
        std::snprintf(tail, sizeof(tail), "\n%d %g", r.mutant.offset, r.score);
        if (done) text += "\n";
        text += mut;
        text += tail;
        done++;
    }
    FILE* o = std::fopen(output_path, "w");
    if (!o) return PSA_ERR_IO;
    bool wrote = std::fwrite(text.data(), 1, text.size(), o) == text.size();
    wrote = (std::fclose(o) == 0) && wrote;
    if (nblocks) *nblocks = done;
    return wrote ? PSA_OK : PSA_ERR_IO;
}

} // extern "C"
