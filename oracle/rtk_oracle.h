/* oracle/rtk_oracle.h — CPU restatement of the reference algorithms (TEST INFRASTRUCTURE ONLY;
 * see rtk_oracle.cpp).  Never included by the product. */
#ifndef RTK_ORACLE_H
#define RTK_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
void* orc_graph_create(int k, uint64_t n, const char* const* unitigs);
void orc_graph_free(void* g);
void orc_free(void* p);
/* hits as 4 x u32 (pos, unitig, dist, strand) in the reference's v_um order; n_lookups = k-mer
 * queries issued by the reference algorithm (findUnitig calls). */
int64_t orc_search_sequence(void* g, const char* s, int exact, int ins, int del, int subst, int or_excl,
                            uint32_t** out, uint64_t* n_lookups);
/* mode 0 NW / 1 SHW / 2 HW; kmax -1 = unbounded; dist -1 if above kmax */
int orc_edit_distance(const char* q, int ql, const char* t, int tl, int mode, int kmax, int iupac, int* dist,
                      int** ends, int* n_ends);
#ifdef __cplusplus
}
#endif
#endif
