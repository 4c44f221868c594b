// renderer_b200.cpp — link this INSTEAD of swegl's src/render/renderer.cpp.
//
// swegl::render (swegl/render/renderer.hpp:27-34, header-only) first runs
// vertex_shader_t::original_to_world(scene) on the CPU -- which also fills every
// node_t::original_to_world_matrix -- and then calls swegl::_render(scene, viewport) per viewport.
// This translation unit supplies that one function, so src/test_1.cpp builds and runs unmodified
// with the frame produced by the CUDA path.  (swegl_b200::render() in swegl_b200_adapter.hpp is the
// faster entry point: it skips the redundant CPU per-vertex loop.)
#include <swegl/render/renderer.hpp>

#include "swegl_b200_adapter.hpp"

namespace swegl
{

void _render(scene_t & scene, viewport_t & viewport)
{
	swegl_b200::engine_t & engine = swegl_b200::default_engine();
	// original_to_world already ran for this frame; begin_frame is cheap (node matrices + lights) and
	// idempotent, so it is simply repeated per viewport
	engine.begin_frame(scene, true);
	engine.render_viewport(scene, viewport);
}

} // namespace swegl
