// psa_host.h -- host-side internals shared by the table resolver, the engine and the C ABI.
#pragma once
#include "psa_b200.h"
#include "psa_common.h"

namespace psa {

// 27-symbol alphabet helpers
int  symbol_index(char c);          // 0..26, -1 outside [A-Z-]
char symbol_char(int idx);
char sign_of(int c1, int c2);       // '*' ':' '.' '_'

// best replacement symbol for the Seq2 symbol c2 facing the Seq1 symbol c1; -1 if none
int best_substitute(int c1, int c2, const double* w, bool is_max);

// resolve the public table and/or the device table; returns PSA_OK / PSA_ERR_WEIGHTS
int build_tables(const double* w, int is_max, long long max_len2, psa_pair_table* pub, DeviceTable* dev);

} // namespace psa
