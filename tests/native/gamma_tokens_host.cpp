// Host build of kmer-db_b200/csrc/gamma_tokens.cuh for the CPU test (tests/test_host.py::test_gamma_tokens_*): the very
// code the decode kernel runs per pattern, compiled by g++, so that its parsing and its row-block stretches can be
// checked against the oracle without a GPU.  Test infrastructure; not part of the product libraries.
#include <cstdint>
#include <vector>

#include "../../kmer-db_b200/csrc/gamma_tokens.cuh"

extern "C" {

// Decodes one list like k_decode_locals does.  out: l ids; closes: triples (row_block, start position, count), at most l.
// Returns 0, or the parser's error (1, 2), or 3 when the deltas sum to more than `last`.
int kdbx_test_decode_list(const uint64_t* w, uint32_t nb, uint32_t l, uint32_t last, int with_blocks, uint32_t sh, uint32_t* out,
                          uint32_t* closes, uint32_t* num_closes, uint32_t* runs_out) {
    *num_closes = 0;
    if (l == 0) return 0;
    if (l == 1) { out[0] = last; *runs_out = 1; if (with_blocks) { closes[0] = last >> sh; closes[1] = 0; closes[2] = 1; *num_closes = 1; } return 0; }
    uint64_t sum = 0;
    uint32_t runs = 1;
    const int rc = gamma_parse_tokens(w, nb, l, out, sum, runs);
    if (rc) return rc;
    if (sum > last) return 3;
    const uint32_t first = last - (uint32_t)sum;
    out[0] = first;
    *runs_out = runs;
    if (!with_blocks) tokens_to_ids(out, l, first);
    else {
        uint32_t n = 0;
        tokens_to_ids_blocks(out, l, first, sh, [&](uint32_t rb, uint32_t j, uint32_t k) { closes[3 * n] = rb; closes[3 * n + 1] = j; closes[3 * n + 2] = k; ++n; });
        *num_closes = n;
    }
    return 0;
}

}
