// swegl_b200_adapter.hpp — C++ host side of the drop-in: swegl's own scene_t / viewport_t objects in,
// C ABI (include/swegl_b200.h) out.  Compiled against the USER'S swegl checkout (it includes swegl's headers,
// it does not copy them); see INTEGRATION.md.
//
//   swegl_b200::render(scene, viewport...)      replaces  swegl::render           swegl/render/renderer.hpp:27-34
//   swegl_b200::engine_t::render_viewport()     replaces  swegl::_render          src/render/renderer.cpp:77-235
//
// Errors: the reference returns void and asserts; here every non-zero C-ABI status becomes a
// std::runtime_error (unknown pixel shader subclass, non-opaque materials with transparency layers, CUDA
// failure).  There is no CPU fallback.
#pragma once

#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <typeinfo>
#include <vector>

#include <swegl/data/model.hpp>
#include <swegl/render/viewport.hpp>
#include <swegl/render/pixel_shaders.hpp>
#include <swegl/render/post_shaders.hpp>

#include "swegl_b200.h"

namespace swegl_b200
{

class engine_t
{
public:
	explicit engine_t(int device = 0)
	{
		if (swegl_b200_create(device, &m_ctx) != SWEGL_B200_OK)
			throw std::runtime_error("swegl_b200: no usable sm_100 CUDA device (there is no CPU fallback)");
	}
	~engine_t() { swegl_b200_destroy(m_ctx); }
	engine_t(const engine_t &) = delete;
	engine_t & operator=(const engine_t &) = delete;

	swegl_b200_ctx * ctx() { return m_ctx; }

	// forget the cached static scene (call after editing vertices / indices / materials / textures)
	void invalidate() { m_signature.clear(); }

	// per frame, once for all viewports: node matrices + lights (vertex_shaders.hpp:16-33 computes the matrices;
	// when `matrices_ready` the caller already ran vertex_shader_t::original_to_world)
	void begin_frame(swegl::scene_t & scene, bool matrices_ready)
	{
		if ( ! matrices_ready)
			for (auto node_idx : scene.root_nodes)
				hierarchy(scene, scene.nodes[node_idx], swegl::matrix44_t::Identity);
		upload_static_if_changed(scene);

		const size_t n = scene.nodes.size();
		m_node_world.resize(16 * n);
		m_node_normal.resize(9 * n);
		for (size_t i = 0; i < n; i++)
		{
			const auto & node = scene.nodes[i];
			for (int r = 0; r < 4; r++)
				for (int c = 0; c < 4; c++)
					m_node_world[16 * i + 4 * r + c] = node.original_to_world_matrix[r][c];
			const swegl::matrix44_t rs = swegl::scale(node.rotation, node.scale);   // vertex_shaders.hpp:63
			for (int r = 0; r < 3; r++)
				for (int c = 0; c < 3; c++)
					m_node_normal[9 * i + 3 * r + c] = rs[r][c];
		}
		m_lights.clear();
		for (const auto & psl : scene.point_source_lights)
		{
			m_lights.push_back(psl.position.x()); m_lights.push_back(psl.position.y());
			m_lights.push_back(psl.position.z()); m_lights.push_back(psl.intensity);
		}
		swegl_b200_frame_desc fd{};
		fd.node_world = m_node_world.data();
		fd.node_normal = m_node_normal.data();
		fd.ambient = scene.ambient_light_intensity;
		fd.sun_dir[0] = scene.sun_direction.x(); fd.sun_dir[1] = scene.sun_direction.y(); fd.sun_dir[2] = scene.sun_direction.z();
		fd.sun_intensity = scene.sun_intensity;
		fd.n_point_lights = (uint32_t)scene.point_source_lights.size();
		fd.point_lights = m_lights.data();
		check(swegl_b200_begin_frame(m_ctx, &fd), "begin_frame");
	}

	// Device-side animation.  set_animation flattens scene.animations, the nodes' TRS as they are now and
	// the hierarchy (children_idx) into swegl_b200_animation_desc; begin_frame_animated(scene, t) then stands for
	//     scene.animate(t);                                   test_1.cpp:378, model.hpp:146-177
	//     vertex_shader_t::original_to_world(scene);          vertex_shaders.hpp:16-33
	// of one frame: the key-frame blend, the TRS -> matrix step and the hierarchy product run on the device, only the time
	// stamp and the lights travel.  The host scene_t is NOT touched (its node_t::rotation / original_to_world_matrix keep
	// their values); mix with begin_frame() at will.
	void set_animation(const swegl::scene_t & scene)
	{
		upload_static_if_changed(scene);
		const size_t n = scene.nodes.size();
		std::vector<int32_t> parent(n, -1);
		std::vector<float> rot(16 * n), tr(3 * n), sc(3 * n), end_time, step_time, step_value;
		for (size_t i = 0; i < n; i++)
		{
			const auto & node = scene.nodes[i];
			for (auto child : node.children_idx)
				if (child >= 0 && (size_t)child < n) parent[(size_t)child] = (int32_t)i;
			for (int r = 0; r < 4; r++)
				for (int c = 0; c < 4; c++)
					rot[16 * i + 4 * r + c] = node.rotation[r][c];
			tr[3 * i] = node.translation.x(); tr[3 * i + 1] = node.translation.y(); tr[3 * i + 2] = node.translation.z();
			sc[3 * i] = node.scale.x(); sc[3 * i + 1] = node.scale.y(); sc[3 * i + 2] = node.scale.z();
		}
		std::vector<swegl_b200_anim_channel> chans;
		for (size_t a = 0; a < scene.animations.size(); a++)
		{
			end_time.push_back(scene.animations[a].end_time);
			for (const auto & ch : scene.animations[a].channels)
			{
				swegl_b200_anim_channel c{};
				c.animation = (int32_t)a; c.node = ch.node_idx; c.path = (int32_t)ch.path;
				c.first_step = (uint32_t)step_time.size(); c.n_steps = (uint32_t)ch.steps.size();
				for (const auto & st : ch.steps)
				{
					step_time.push_back(st.time);
					step_value.push_back(st.value.x()); step_value.push_back(st.value.y());
					step_value.push_back(st.value.z()); step_value.push_back(st.value.w());
				}
				chans.push_back(c);
			}
		}
		swegl_b200_animation_desc ad{};
		ad.n_nodes = (uint32_t)n;
		ad.node_parent = parent.data(); ad.node_rotation = rot.data(); ad.node_translation = tr.data(); ad.node_scale = sc.data();
		ad.n_animations = (uint32_t)end_time.size(); ad.end_time = end_time.data();
		ad.n_channels = (uint32_t)chans.size(); ad.channels = chans.data();
		ad.n_steps = (uint32_t)step_time.size(); ad.step_time = step_time.data(); ad.step_value = step_value.data();
		check(swegl_b200_set_animation(m_ctx, &ad), "set_animation");
		m_anim_uploaded = true;
	}

	void begin_frame_animated(const swegl::scene_t & scene, float elapsed_seconds)
	{
		upload_static_if_changed(scene);                // (a re-upload drops the device's animation tables)
		if ( ! m_anim_uploaded) set_animation(scene);
		m_lights.clear();
		for (const auto & psl : scene.point_source_lights)
		{
			m_lights.push_back(psl.position.x()); m_lights.push_back(psl.position.y());
			m_lights.push_back(psl.position.z()); m_lights.push_back(psl.intensity);
		}
		swegl_b200_frame_desc fd{};
		fd.ambient = scene.ambient_light_intensity;
		fd.sun_dir[0] = scene.sun_direction.x(); fd.sun_dir[1] = scene.sun_direction.y(); fd.sun_dir[2] = scene.sun_direction.z();
		fd.sun_intensity = scene.sun_intensity;
		fd.n_point_lights = (uint32_t)scene.point_source_lights.size();
		fd.point_lights = m_lights.data();
		check(swegl_b200_begin_frame_animated(m_ctx, elapsed_seconds, &fd), "begin_frame_animated");
	}

	// the device screen follows the SDL surface the viewports draw into
	void ensure_screen(const SDL_Surface * screen)
	{
		if (screen->w != m_screen_w || screen->h != m_screen_h)
		{
			check(swegl_b200_set_screen(m_ctx, screen->w, screen->h), "set_screen");
			m_screen_w = screen->w; m_screen_h = screen->h;
		}
	}

	// viewport_t + camera_t + shader selection -> the C ABI's description of one view
	static swegl_b200_viewport_desc describe(swegl::viewport_t & vp)
	{
		swegl_b200_viewport_desc d{};
		d.x = vp.m_x; d.y = vp.m_y; d.w = vp.m_w; d.h = vp.m_h;
		for (int r = 0; r < 4; r++)
			for (int c = 0; c < 4; c++)
			{
				d.view[4 * r + c] = vp.camera().m_viewmatrix[r][c];
				d.proj[4 * r + c] = vp.camera().m_projectionmatrix[r][c];
			}
		const swegl::vertex_t cam = vp.camera().position();
		d.cam_pos[0] = cam.x(); d.cam_pos[1] = cam.y(); d.cam_pos[2] = cam.z();
		d.vp_m00 = vp.m_viewportmatrix[0][0]; d.vp_m03 = vp.m_viewportmatrix[0][3];
		d.vp_m11 = vp.m_viewportmatrix[1][1]; d.vp_m13 = vp.m_viewportmatrix[1][3];
		select_shader(*vp.m_pixel_shader, d);
		d.post_mode = SWEGL_B200_POST_NULL;
		if (auto * dof = dynamic_cast<swegl::post_shader_depth_box *>(vp.m_post_shader))
		{
			d.post_mode = SWEGL_B200_POST_DOF;          // the repaired DoF-R semantics, see DESIGN.md
			d.focal_distance = dof->focal_distance;
			d.focal_depth = dof->focal_depth;
		}
		d.transparency_layers = vp.m_got_transparency ? (int32_t)vp.m_transparency_layers.size() : 0;
		return d;
	}

	// the body of swegl::_render(scene, viewport): the frame is in vp.m_screen->pixels / vp.zbuffer() on return
	void render_viewport(swegl::scene_t &, swegl::viewport_t & vp)
	{
		SDL_Surface * screen = vp.m_screen;
		ensure_screen(screen);
		const swegl_b200_viewport_desc d = describe(vp);
		check(swegl_b200_render_viewport(m_ctx, &d, screen->pixels, screen->pitch, vp.zbuffer(), nullptr), "render_viewport");
	}

	// the same with the result left in HBM (asynchronous when stats == nullptr): the multi-context hosts of
	// swegl_b200_host.hpp read it back themselves
	void render_viewport_device(const swegl_b200_viewport_desc & d, swegl_b200_stats * stats = nullptr)
	{
		check(swegl_b200_render_viewport_device(m_ctx, &d, stats), "render_viewport_device");
	}

	void check(int rc, const char * what)
	{
		if (rc != SWEGL_B200_OK)
			throw std::runtime_error(std::string("swegl_b200 ") + what + ": status " + std::to_string(rc) + ": " + swegl_b200_last_error(m_ctx));
	}

private:
	swegl_b200_ctx * m_ctx = nullptr;
	int m_screen_w = 0, m_screen_h = 0;
	std::string m_signature;
	bool m_anim_uploaded = false;
	std::vector<float> m_node_world, m_node_normal, m_lights;

	// node_t::original_to_world_matrix for the whole hierarchy (vertex_shaders.hpp:16-18,26-27), without the
	// per-vertex loop: the device computes v_world
	static void hierarchy(swegl::scene_t & scene, swegl::node_t & node, const swegl::matrix44_t & parent)
	{
		node.original_to_world_matrix = parent * node.get_local_world_matrix();
		for (auto child_idx : node.children_idx)
			hierarchy(scene, scene.nodes[child_idx], node.original_to_world_matrix);
	}

	template <typename L>
	static bool match_light_and_texture(const swegl::pixel_shader_t & ps, int light, swegl_b200_viewport_desc & d)
	{
		using namespace swegl;
		const std::type_info & t = typeid(ps);
		int tex = -1;
		if (t == typeid(pixel_shader_light_and_texture<L, pixel_shader_t>)) tex = SWEGL_B200_TEX_PLAIN;
		else if (t == typeid(pixel_shader_light_and_texture<L, pixel_shader_texture>)) tex = SWEGL_B200_TEX_NEAREST;
		else if (t == typeid(pixel_shader_light_and_texture<L, pixel_shader_texture_bilinear>)) tex = SWEGL_B200_TEX_BILINEAR;
		if (tex < 0) return false;
		d.light_mode = light; d.tex_mode = tex;
		return true;
	}

	// the built-in pixel_shader_t family (pixel_shaders.hpp:15-179); user subclasses cannot run on the device
	static void select_shader(const swegl::pixel_shader_t & ps, swegl_b200_viewport_desc & d)
	{
		using namespace swegl;
		if (match_light_and_texture<pixel_shader_lights_phong>(ps, SWEGL_B200_LIGHT_PHONG, d)) return;
		if (match_light_and_texture<pixel_shader_lights_flat>(ps, SWEGL_B200_LIGHT_FLAT, d)) return;
		const std::type_info & t = typeid(ps);
		d.light_mode = SWEGL_B200_LIGHT_NONE;
		if (t == typeid(pixel_shader_t)) { d.tex_mode = SWEGL_B200_TEX_PLAIN; return; }
		if (t == typeid(pixel_shader_texture)) { d.tex_mode = SWEGL_B200_TEX_NEAREST; return; }
		if (t == typeid(pixel_shader_texture_bilinear)) { d.tex_mode = SWEGL_B200_TEX_BILINEAR; return; }
		throw std::runtime_error(std::string("swegl_b200: pixel shader type not available on the device: ") + t.name());
	}

	void upload_static_if_changed(const swegl::scene_t & scene)
	{
		// cheap identity of the static part: counts and buffer addresses
		std::string sig;
		auto add = [&sig](const void * p, size_t n) { sig.append(reinterpret_cast<const char *>(&p), sizeof p); sig.append(reinterpret_cast<const char *>(&n), sizeof n); };
		for (const auto & node : scene.nodes)
			for (const auto & prim : node.primitives)
			{
				add(prim.vertices.data(), prim.vertices.size());
				add(prim.indices.data(), prim.indices.size());
			}
		add(scene.materials.data(), scene.materials.size());
		for (const auto & img : scene.images)
			add(img.m_mipmaps.empty() ? nullptr : img.m_mipmaps[0]->m_bitmap, img.m_mipmaps.size());
		if (sig == m_signature) return;

		std::vector<swegl_b200_primitive> prims;
		std::vector<float> pos, nrm, uv;
		std::vector<uint32_t> idx;
		for (size_t n = 0; n < scene.nodes.size(); n++)
			for (const auto & prim : scene.nodes[n].primitives)
			{
				swegl_b200_primitive p{};
				p.node = (int32_t)n; p.mode = (int32_t)prim.mode; p.material_id = prim.material_id;
				p.first_vertex = (uint32_t)(pos.size() / 3); p.n_vertices = (uint32_t)prim.vertices.size();
				p.first_index = (uint32_t)idx.size(); p.n_indices = (uint32_t)prim.indices.size();
				for (const auto & mv : prim.vertices)
				{
					pos.push_back(mv.v.x()); pos.push_back(mv.v.y()); pos.push_back(mv.v.z());
					nrm.push_back(mv.normal.x()); nrm.push_back(mv.normal.y()); nrm.push_back(mv.normal.z());
					uv.push_back(mv.tex_coords.x()); uv.push_back(mv.tex_coords.y());
				}
				idx.insert(idx.end(), prim.indices.begin(), prim.indices.end());
				prims.push_back(p);
			}
		auto mat = [](const swegl::material_t & m) {
			swegl_b200_material o{};
			o.b = m.color.o.b; o.g = m.color.o.g; o.r = m.color.o.r; o.a = m.color.o.a;
			o.metallic = m.metallic; o.roughness = m.roughness; o.texture_idx = m.texture_idx; o.double_sided = m.double_sided ? 1 : 0;
			return o;
		};
		std::vector<swegl_b200_material> mats;
		for (const auto & m : scene.materials) mats.push_back(mat(m));
		std::vector<swegl_b200_texture> texs;
		for (const auto & img : scene.images)
		{
			swegl_b200_texture t{};
			t.texels = img.m_mipmaps[0]->m_bitmap;                      // level 0 only (pixel_shaders.cpp:298-300)
			t.width = (int32_t)img.m_mipmaps[0]->m_width; t.height = (int32_t)img.m_mipmaps[0]->m_height;
			texs.push_back(t);
		}
		swegl_b200_scene_desc sd{};
		sd.n_nodes = (uint32_t)scene.nodes.size(); sd.n_primitives = (uint32_t)prims.size();
		sd.n_vertices = (uint32_t)(pos.size() / 3); sd.n_indices = (uint32_t)idx.size();
		sd.n_materials = (uint32_t)mats.size(); sd.n_textures = (uint32_t)texs.size();
		sd.primitives = prims.data(); sd.positions = pos.data(); sd.normals = nrm.data(); sd.texcoords = uv.data();
		sd.indices = idx.data(); sd.materials = mats.data(); sd.default_material = mat(scene.default_material);
		sd.textures = texs.data();
		check(swegl_b200_upload_scene(m_ctx, &sd), "upload_scene");
		m_signature = sig;
		m_anim_uploaded = false;
	}
};

// process-wide engine used by the replacement swegl::_render (renderer_b200.cpp)
inline engine_t & default_engine()
{
	static engine_t engine(0);
	return engine;
}

// drop-in for swegl::render(scene, viewports...): the per-vertex original_to_world loop runs on the device
template <typename... T>
void render(swegl::scene_t & scene, T &... viewports)
{
	engine_t & e = default_engine();
	e.begin_frame(scene, false);
	(e.render_viewport(scene, viewports), ...);
}

// drop-in for the application's pair  scene.animate(t); swegl::render(scene, viewports...)  (test_1.cpp:374-378) with the
// animation evaluated on the device
template <typename... T>
void render_animated(swegl::scene_t & scene, float elapsed_seconds, T &... viewports)
{
	engine_t & e = default_engine();
	e.begin_frame_animated(scene, elapsed_seconds);
	(e.render_viewport(scene, viewports), ...);
}

} // namespace swegl_b200
