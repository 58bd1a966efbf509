// xo_clcompat_mcml.cuh -- layer-stack accessors of the layered simulator for
// user-written plugin fragments (mcml.template.h:642-764, mclayer/layer.py:
// 105-170).  Included after the plugin slots are bound (XoPf ...), before the
// fragments' implementations.
#pragma once
#include "xo_clcompat.cuh"
#include "mcml_layer.cuh"

typedef xo::MlLayer McLayer;
#define mcsim_layer(psim, index) (static_cast<const McLayer *>((psim)->layers) + (index))
#define mcsim_current_layer(psim) mcsim_layer(psim, (psim)->state.layer_index)
#define mcsim_top_layer_index(psim) (0)
#define mcsim_top_layer(psim) mcsim_layer(psim, 0)
#define mcsim_top_sample_layer_index(psim) (1)
#define mcsim_top_sample_layer(psim) mcsim_layer(psim, 1)
#define mcsim_bottom_layer_index(psim) (mcsim_layer_count(psim) - 1)
#define mcsim_bottom_layer(psim) mcsim_layer(psim, mcsim_layer_count(psim) - 1)
#define mcsim_bottom_sample_layer_index(psim) (mcsim_layer_count(psim) - 2)
#define mcsim_bottom_sample_layer(psim) mcsim_layer(psim, mcsim_layer_count(psim) - 2)
#define mc_layer_thickness(player) ((player)->thickness)
#define mc_layer_top(player) ((player)->top)
#define mc_layer_bottom(player) ((player)->bottom)
#define mc_layer_n(player) ((player)->n)
#define mc_layer_cc_top(player) ((player)->cc_top)
#define mc_layer_cc_bottom(player) ((player)->cc_bottom)
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
