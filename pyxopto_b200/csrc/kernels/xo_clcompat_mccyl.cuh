// xo_clcompat_mccyl.cuh -- layer accessors of the cylindrical simulator for user-written
// plugin fragments (mccyl.template.h:655-762, mccyl/mclayer/layer.py:167-258).  Included
// after the plugin slots are bound (XoPf ...), before the fragments' implementations.
#pragma once
#include "xo_clcompat.cuh"
#include "mccyl_layer.cuh"

typedef xo::CylLayer McLayer;
#define mcsim_layer(psim, index) (static_cast<const McLayer *>((psim)->layers) + (index))
#define mcsim_current_layer(psim) mcsim_layer(psim, (psim)->state.layer_index)
#define mcsim_layer_index_is_sample(psim, index) ((index) > 0 && (index) < mcsim_layer_count(psim))
#define mcsim_top_layer_index(psim) (0)
#define mcsim_top_layer(psim) mcsim_layer(psim, 0)
#define mcsim_top_sample_layer_index(psim) (1)
#define mcsim_top_sample_layer(psim) mcsim_layer(psim, 1)
#define mcsim_bottom_layer_index(psim) (mcsim_layer_count(psim) - 1)
#define mcsim_bottom_layer(psim) mcsim_layer(psim, mcsim_layer_count(psim) - 1)
#define mc_layer_r_inner(player) ((player)->r_inner)
#define mc_layer_r_outer(player) ((player)->r_outer)
#define mc_layer_n(player) ((player)->n)
#define mc_layer_cc_inner(player) ((player)->cc_inner)
#define mc_layer_cc_outer(player) ((player)->cc_outer)
#if XO_ANISO
#define mc_layer_mus(player, pdir) ((player)->mus_at(*(pdir)))
#define mc_layer_mua(player, pdir) ((player)->mua_at(*(pdir)))
#define mc_layer_mut(player, pdir) (xo::tensor_project((player)->mut_t, *(pdir)))
#define mc_layer_inv_mut(player, pdir) ((player)->inv_mut_at(*(pdir)))
#define mc_layer_mua_inv_mut(player, pdir) ((player)->mua_inv_mut_at(*(pdir)))
#else
#define mc_layer_mus(player, ...) ((player)->mus)
#define mc_layer_mua(player, ...) ((player)->mua)
#define mc_layer_mut(player, ...) ((player)->mua + (player)->mus)
#define mc_layer_inv_mut(player, ...) ((player)->inv_mut)
#define mc_layer_mua_inv_mut(player, ...) ((player)->mua_inv_mut)
#endif
