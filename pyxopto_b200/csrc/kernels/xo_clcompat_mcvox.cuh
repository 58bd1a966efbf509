// xo_clcompat_mcvox.cuh -- voxel-geometry accessors of the voxelised simulator for
// user-written plugin fragments (mcvox.template.h:482-930, mcbase/mcmaterial.py:74-140).
// Included after the plugin slots are bound (XoPf ...), before the fragments'
// implementations.  The kernel locates the voxel of a launched packet from its position
// (mcvox.template.c:259-262, mcsim_launch_one): the voxel-index setters only update the
// facade.
#pragma once
#include "xo_clcompat.cuh"
#include "mcvox_medium.cuh"

typedef xo::VoxMaterial McMaterial;
typedef xo::VoxCfg McVoxelConfig;
#define __mc_material_mem
#define __mc_geometry_mem
#define mcsim_voxel_config(psim) (static_cast<const McVoxelConfig *>((psim)->voxel_cfg))
#define mcsim_top_left(psim) (&mcsim_voxel_config(psim)->top_left)
#define mcsim_top_left_x(psim) (mcsim_voxel_config(psim)->top_left.x)
#define mcsim_top_left_y(psim) (mcsim_voxel_config(psim)->top_left.y)
#define mcsim_top_left_z(psim) (mcsim_voxel_config(psim)->top_left.z)
#define mcsim_bottom_right(psim) (&mcsim_voxel_config(psim)->bottom_right)
#define mcsim_bottom_right_x(psim) (mcsim_voxel_config(psim)->bottom_right.x)
#define mcsim_bottom_right_y(psim) (mcsim_voxel_config(psim)->bottom_right.y)
#define mcsim_bottom_right_z(psim) (mcsim_voxel_config(psim)->bottom_right.z)
#define mcsim_voxel_size(psim) (&mcsim_voxel_config(psim)->size)
#define mcsim_voxel_size_x(psim) (mcsim_voxel_config(psim)->size.x)
#define mcsim_voxel_size_y(psim) (mcsim_voxel_config(psim)->size.y)
#define mcsim_voxel_size_z(psim) (mcsim_voxel_config(psim)->size.z)
#define mcsim_shape_x(psim) (mcsim_voxel_config(psim)->nx)
#define mcsim_shape_y(psim) (mcsim_voxel_config(psim)->ny)
#define mcsim_shape_z(psim) (mcsim_voxel_config(psim)->nz)
#define mcsim_material(psim, index) (static_cast<const McMaterial *>((psim)->materials) + (index))
#define mcsim_surrounding_material(psim) mcsim_material(psim, 0)
#define mcsim_surrounding_material_index(psim) (0)
#define mcsim_voxels(psim) ((psim)->voxels)
#define mcsim_flat_voxel_index(psim, pindex) \
	(((pindex)->z*mcsim_shape_y(psim) + (pindex)->y)*mcsim_shape_x(psim) + (pindex)->x)
#define mcsim_voxel_material_index(psim, pindex) \
	(mcsim_voxels(psim)[mcsim_flat_voxel_index(psim, pindex)])
#define mcsim_voxel_material(psim, pindex) \
	mcsim_material(psim, mcsim_voxel_material_index(psim, pindex))
#define mcsim_voxel_index(psim) (&(psim)->state.voxel_index)
#define mcsim_voxel_index_x(psim) ((psim)->state.voxel_index.x)
#define mcsim_voxel_index_y(psim) ((psim)->state.voxel_index.y)
#define mcsim_voxel_index_z(psim) ((psim)->state.voxel_index.z)
#define mcsim_current_voxel_material_index(psim) ((psim)->state.voxel_material_index)
#define mcsim_current_voxel_material(psim) \
	mcsim_material(psim, mcsim_current_voxel_material_index(psim))
#define mcsim_set_voxel_index_components(psim, x_, y_, z_) \
	{ (psim)->state.voxel_index.x = (x_); (psim)->state.voxel_index.y = (y_); (psim)->state.voxel_index.z = (z_); }
#define mcsim_set_voxel_index(psim, pindex) \
	mcsim_set_voxel_index_components(psim, (pindex)->x, (pindex)->y, (pindex)->z)
#define mcsim_set_voxel_material_index(psim, material_index) \
	{ (psim)->state.voxel_material_index = (material_index); }
#define mcsim_set_voxel_index_components_and_material(psim, x_, y_, z_) \
	{ mcsim_set_voxel_index_components(psim, x_, y_, z_); \
	  (psim)->state.voxel_material_index = mcsim_voxel_material_index(psim, &(psim)->state.voxel_index); }
#define mcsim_set_voxel_index_and_material(psim, pindex) \
	mcsim_set_voxel_index_components_and_material(psim, (pindex)->x, (pindex)->y, (pindex)->z)
#define mc_material_n(pmaterial) ((pmaterial)->n)
#define mc_material_pf(pmaterial) (&(pmaterial)->pf)
#if XO_ANISO
#define mc_material_mus(pmaterial, pdir) ((pmaterial)->mus_at(*(pdir)))
#define mc_material_mua(pmaterial, pdir) ((pmaterial)->mua_at(*(pdir)))
#define mc_material_mut(pmaterial, pdir) (xo::tensor_project((pmaterial)->mut_t, *(pdir)))
#define mc_material_inv_mut(pmaterial, pdir) ((pmaterial)->inv_mut_at(*(pdir)))
#define mc_material_mua_inv_mut(pmaterial, pdir) ((pmaterial)->mua_inv_mut_at(*(pdir)))
#else
#define mc_material_mus(pmaterial, ...) ((pmaterial)->mus)
#define mc_material_mua(pmaterial, ...) ((pmaterial)->mua)
#define mc_material_mut(pmaterial, ...) ((pmaterial)->mua + (pmaterial)->mus)
#define mc_material_inv_mut(pmaterial, ...) ((pmaterial)->inv_mut)
#define mc_material_mua_inv_mut(pmaterial, ...) ((pmaterial)->mua_inv_mut)
#endif
