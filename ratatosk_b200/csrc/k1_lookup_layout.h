// k1_lookup_layout.h — bit layout of the labelled raw hits the K1 kernels emit (host + device).
#pragma once
#include <stdint.h>
// hit.a = read(24) | variant(10) | pos_s(30) ; hit.b = P(40) | strand<<40
// variant order index = the order searchSequence runs its passes in (Search.tcc:717-765):
//   subst: shift*4+letter in [0,4k) ; ins: 4k + shift*4+letter ; del: 8k + shift
#define RTK_HIT_POS_BITS 30
#define RTK_HIT_VAR_BITS 10
#define RTK_HIT_READ_BITS 24
