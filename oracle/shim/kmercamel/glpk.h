// Shim: GLPK is absent from this image. kmercamel reaches it only from `optimize -a runs`
// (masks.h:121-164), which the data generators never use. Every entry point aborts.
#pragma once
#include <cstdlib>
struct glp_prob;
enum { GLP_MIN = 1, GLP_LO = 2, GLP_IV = 2, GLP_OFF = 0 };
[[noreturn]] inline void glp_shim_abort_() { std::abort(); }
inline glp_prob* glp_create_prob() { glp_shim_abort_(); }
inline void glp_set_obj_dir(glp_prob*, int) { glp_shim_abort_(); }
inline int glp_add_rows(glp_prob*, int) { glp_shim_abort_(); }
inline int glp_add_cols(glp_prob*, int) { glp_shim_abort_(); }
inline void glp_set_row_bnds(glp_prob*, int, int, double, double) { glp_shim_abort_(); }
inline void glp_set_col_bnds(glp_prob*, int, int, double, double) { glp_shim_abort_(); }
inline void glp_set_col_kind(glp_prob*, int, int) { glp_shim_abort_(); }
inline void glp_set_obj_coef(glp_prob*, int, double) { glp_shim_abort_(); }
inline int glp_term_out(int) { glp_shim_abort_(); }
inline void glp_load_matrix(glp_prob*, int, const int*, const int*, const double*) { glp_shim_abort_(); }
inline int glp_simplex(glp_prob*, const void*) { glp_shim_abort_(); }
inline double glp_get_col_prim(glp_prob*, int) { glp_shim_abort_(); }
inline void glp_delete_prob(glp_prob*) { glp_shim_abort_(); }
