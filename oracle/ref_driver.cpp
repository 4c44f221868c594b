// TEST INFRASTRUCTURE ONLY — C driver around the UNMODIFIED reference (gbizzotto/swegl).
//
// oracle/Makefile compiles this file together with the reference's own
// src/{render,projection,data}/*.cpp (where they lie under /root/reference) and the shim
// headers in oracle/shims/ into oracle/_ref/libswegl_ref.so.  Nothing here re-implements the
// renderer: it loads / builds a swegl::scene_t, sets up a swegl::viewport_t exactly like
// src/test_1.cpp:328-361 and calls swegl::render(), then exposes the state the reference
// leaves behind (SDL_Surface::pixels, viewport_t::m_zbuffer, mesh_vertex_t fields) so the
// restatement in swegl_oracle.c and the CUDA path can be compared with it bit for bit.
// It is also the "reference" CPU baseline timed by bench.py.
//
// The image decoders (src/misc/image.cpp needs libpng/libjpeg, absent here) are replaced by
// a callback so the caller decodes the embedded PNG/JPEG once (PIL) and every party samples
// the same texels.

#include <chrono>
#include <cstring>
#include <memory>
#include <vector>

#include <swegl/data/gltf.hpp>
#include <swegl/data/model.hpp>
#include <swegl/render/renderer.hpp>
#include <swegl/render/vertex_shaders.hpp>
#include <swegl/render/pixel_shaders.hpp>
#include <swegl/render/post_shaders.hpp>

typedef int (*ref_decode_cb)(const char * filename, int offset, int * w, int * h, unsigned int * out_bgra);
static ref_decode_cb g_decode = nullptr;

namespace swegl
{
// replacements for src/misc/image.cpp:18-258 (decode delegated to the caller)
static texture_t decode_via_callback(const std::string & filename, int offset)
{
	int w = 0, h = 0;
	if (!g_decode || g_decode(filename.c_str(), offset, &w, &h, nullptr) != 0 || w <= 0 || h <= 0)
		return texture_t(nullptr, 0, 0);
	unsigned int * data = new unsigned int[(size_t)w * h];
	if (g_decode(filename.c_str(), offset, &w, &h, data) != 0)
	{
		delete[] data;
		return texture_t(nullptr, 0, 0);
	}
	return texture_t(data, w, h);
}
texture_t read_image_file(const std::string & filename, int offset) { return decode_via_callback(filename, offset); }
texture_t read_png_file(const std::string & filename, int offset) { return decode_via_callback(filename, offset); }
texture_t read_jpeg_file(const std::string & filename, int offset) { return decode_via_callback(filename, offset); }
texture_t read_bmp_file(const std::string & filename, int offset) { return decode_via_callback(filename, offset); }
} // namespace swegl

namespace
{
struct ref_scene
{
	swegl::scene_t scene;
	int spare = 0; // trailing clip slots per primitive that are not real vertices (gltf.cpp:175)
};

struct ref_screen
{
	SDL_PixelFormat format{4};
	SDL_Surface surface{};
	std::vector<unsigned int> pixels;
};

struct ref_viewport
{
	ref_screen * screen;
	std::shared_ptr<swegl::pixel_shader_t> shader;
	std::unique_ptr<swegl::viewport_t> vp;
	swegl::post_shader_t post_null;
	std::unique_ptr<swegl::post_shader_depth_box> post_dof;     // installed by ref_viewport_set_dof
};

template <typename L>
std::shared_ptr<swegl::pixel_shader_t> make_combined(int tex_mode)
{
	using namespace swegl;
	switch (tex_mode)
	{
		case 0: return std::make_shared<pixel_shader_light_and_texture<L, pixel_shader_t>>();
		case 1: return std::make_shared<pixel_shader_light_and_texture<L, pixel_shader_texture>>();
		case 2: return std::make_shared<pixel_shader_light_and_texture<L, pixel_shader_texture_bilinear>>();
	}
	return nullptr;
}

void reserve_clip_slots(swegl::scene_t & scene)
{
	// src/test_1.cpp:337-339
	for (auto & node : scene.nodes)
		for (auto & primitive : node.primitives)
			primitive.vertices.reserve(primitive.vertices.size() + 2);
}
} // namespace

#ifdef SWEGL_B200_DROPIN
#include "swegl_b200_host.hpp"
#endif

extern "C"
{

void ref_set_image_decoder(ref_decode_cb cb) { g_decode = cb; }

void * ref_scene_load(const char * path)
{
	auto * s = new ref_scene;
	s->scene = swegl::load_scene(path);
	s->spare = 2;
	reserve_clip_slots(s->scene);
	return s;
}

void * ref_scene_new() { return new ref_scene; }
void ref_scene_free(void * h) { delete static_cast<ref_scene *>(h); }

int ref_scene_add_material(void * h, int b, int g, int r, int a, float metallic, float roughness, int texture_idx, int double_sided)
{
	auto & sc = static_cast<ref_scene *>(h)->scene;
	sc.materials.push_back(swegl::material_t{swegl::pixel_colors((unsigned char)b, (unsigned char)g, (unsigned char)r, (unsigned char)a),
	                                         metallic, roughness, texture_idx, double_sided != 0});
	return (int)sc.materials.size() - 1;
}

int ref_scene_add_texture(void * h, const unsigned int * bgra, int w, int hgt)
{
	auto & sc = static_cast<ref_scene *>(h)->scene;
	unsigned int * data = new unsigned int[(size_t)w * hgt];
	std::memcpy(data, bgra, sizeof(unsigned int) * (size_t)w * hgt);
	sc.images.emplace_back(data, w, hgt);
	return (int)sc.images.size() - 1;
}

// kind: 0 make_tri(size) 1 make_cube(size) 2 make_tore(precision) 3 make_sphere(precision, size=radius)
// (swegl/data/model.hpp:233-464).  rot = 3 Euler angles applied as rotate_x, rotate_y, rotate_z.
int ref_scene_add_builtin(void * h, int kind, unsigned precision, float size, int material_idx,
                          const float * scale3, const float * rot3, const float * trans3)
{
	auto & sc = static_cast<ref_scene *>(h)->scene;
	swegl::node_t node;
	switch (kind)
	{
		case 0: node = swegl::make_tri(size, material_idx); break;
		case 1: node = swegl::make_cube(size, material_idx); break;
		case 2: node = swegl::make_tore(precision, material_idx); break;
		case 3: node = swegl::make_sphere(precision, size, material_idx); break;
		default: return -1;
	}
	node.rotation = swegl::matrix44_t::Identity;
	if (rot3)
	{
		if (rot3[0] != 0) node.rotation.rotate_x(rot3[0]);
		if (rot3[1] != 0) node.rotation.rotate_y(rot3[1]);
		if (rot3[2] != 0) node.rotation.rotate_z(rot3[2]);
	}
	if (scale3) node.scale = swegl::vertex_t(scale3[0], scale3[1], scale3[2]);
	if (trans3) node.translation = swegl::vertex_t(trans3[0], trans3[1], trans3[2]);
	sc.nodes.emplace_back(std::move(node));
	sc.root_nodes.push_back((int)sc.nodes.size() - 1);
	reserve_clip_slots(sc);
	return (int)sc.nodes.size() - 1;
}

// build a scene_t from flattened arrays (a scene pack): the inverse of ref_scene_export
void * ref_scene_import(unsigned n_nodes, const float * node_scale, const float * node_rotation, const float * node_translation,
                        const int * node_parent,
                        unsigned n_prims, const int * prim_node, const int * prim_mode, const int * prim_material,
                        const unsigned * prim_first_vertex, const unsigned * prim_n_vertices,
                        const unsigned * prim_first_index, const unsigned * prim_n_indices,
                        const float * positions, const float * normals, const float * texcoords, const unsigned * indices)
{
	auto * s = new ref_scene;
	auto & sc = s->scene;
	sc.nodes.resize(n_nodes);
	for (unsigned n = 0; n < n_nodes; n++)
	{
		auto & node = sc.nodes[n];
		node.scale = swegl::vertex_t(node_scale[3 * n], node_scale[3 * n + 1], node_scale[3 * n + 2]);
		for (int r = 0; r < 4; r++)
			for (int c = 0; c < 4; c++)
				node.rotation[r][c] = node_rotation[16 * n + 4 * r + c];
		node.translation = swegl::vertex_t(node_translation[3 * n], node_translation[3 * n + 1], node_translation[3 * n + 2]);
		node.root = node_parent[n] < 0;
	}
	for (unsigned n = 0; n < n_nodes; n++)
	{
		if (node_parent[n] >= 0) sc.nodes[node_parent[n]].children_idx.push_back((int)n);
		else sc.root_nodes.push_back((int)n);
	}
	for (unsigned p = 0; p < n_prims; p++)
	{
		auto & prim = sc.nodes[prim_node[p]].primitives.emplace_back();
		prim.mode = (swegl::primitive_t::index_mode_t)prim_mode[p];
		prim.material_id = prim_material[p];
		prim.vertices.resize(prim_n_vertices[p]);
		for (unsigned k = 0; k < prim_n_vertices[p]; k++)
		{
			unsigned i = prim_first_vertex[p] + k;
			auto & mv = prim.vertices[k];
			mv.v = swegl::vertex_t(positions[3 * i], positions[3 * i + 1], positions[3 * i + 2]);
			// plain member stores: the stored normals must not be re-normalised (gltf.cpp:192-194)
			mv.normal.x() = normals[3 * i]; mv.normal.y() = normals[3 * i + 1]; mv.normal.z() = normals[3 * i + 2];
			mv.tex_coords = swegl::vec2f_t(texcoords[2 * i], texcoords[2 * i + 1]);
		}
		prim.indices.assign(indices + prim_first_index[p], indices + prim_first_index[p] + prim_n_indices[p]);
	}
	reserve_clip_slots(sc);
	return s;
}

void ref_scene_set_lights(void * h, float ambient, float sx, float sy, float sz, float sun_intensity,
                          unsigned n_lights, const float * lights4)
{
	auto & sc = static_cast<ref_scene *>(h)->scene;
	sc.ambient_light_intensity = ambient;
	sc.sun_direction = swegl::normal_t(sx, sy, sz); // normalising ctor, as src/test_1.cpp:335
	sc.sun_intensity = sun_intensity;
	sc.point_source_lights.clear();
	for (unsigned i = 0; i < n_lights; i++)
		sc.point_source_lights.emplace_back(swegl::point_source_light{{lights4[4 * i], lights4[4 * i + 1], lights4[4 * i + 2]}, lights4[4 * i + 3]});
}

void ref_scene_get_sun(void * h, float * out3)
{
	auto & sc = static_cast<ref_scene *>(h)->scene;
	out3[0] = sc.sun_direction.x(); out3[1] = sc.sun_direction.y(); out3[2] = sc.sun_direction.z();
}

// counts: nodes, primitives, vertices (real), indices, materials, textures
void ref_scene_counts(void * h, unsigned * out6)
{
	auto * s = static_cast<ref_scene *>(h);
	unsigned np = 0, nv = 0, ni = 0;
	for (auto & node : s->scene.nodes)
		for (auto & prim : node.primitives)
		{
			np++;
			nv += (unsigned)prim.vertices.size() - s->spare;
			ni += (unsigned)prim.indices.size();
		}
	out6[0] = (unsigned)s->scene.nodes.size(); out6[1] = np; out6[2] = nv; out6[3] = ni;
	out6[4] = (unsigned)s->scene.materials.size(); out6[5] = (unsigned)s->scene.images.size();
}

void ref_scene_export(void * h, float * node_scale, float * node_rotation, float * node_translation, int * node_parent,
                      int * prim_node, int * prim_mode, int * prim_material,
                      unsigned * prim_first_vertex, unsigned * prim_n_vertices, unsigned * prim_first_index, unsigned * prim_n_indices,
                      float * positions, float * normals, float * texcoords, unsigned * indices,
                      unsigned char * mat_bgra, float * mat_metal_rough, int * mat_tex_ds, int * tex_wh)
{
	auto * s = static_cast<ref_scene *>(h);
	auto & sc = s->scene;
	for (size_t n = 0; n < sc.nodes.size(); n++) node_parent[n] = -1;
	unsigned p = 0, v = 0, ix = 0;
	for (size_t n = 0; n < sc.nodes.size(); n++)
	{
		auto & node = sc.nodes[n];
		node_scale[3 * n] = node.scale.x(); node_scale[3 * n + 1] = node.scale.y(); node_scale[3 * n + 2] = node.scale.z();
		for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) node_rotation[16 * n + 4 * r + c] = node.rotation[r][c];
		node_translation[3 * n] = node.translation.x(); node_translation[3 * n + 1] = node.translation.y(); node_translation[3 * n + 2] = node.translation.z();
		for (int child : node.children_idx) node_parent[child] = (int)n;
		for (auto & prim : node.primitives)
		{
			unsigned nv = (unsigned)prim.vertices.size() - s->spare;
			prim_node[p] = (int)n; prim_mode[p] = (int)prim.mode; prim_material[p] = prim.material_id;
			prim_first_vertex[p] = v; prim_n_vertices[p] = nv; prim_first_index[p] = ix; prim_n_indices[p] = (unsigned)prim.indices.size();
			for (unsigned k = 0; k < nv; k++, v++)
			{
				auto & mv = prim.vertices[k];
				positions[3 * v] = mv.v.x(); positions[3 * v + 1] = mv.v.y(); positions[3 * v + 2] = mv.v.z();
				normals[3 * v] = mv.normal.x(); normals[3 * v + 1] = mv.normal.y(); normals[3 * v + 2] = mv.normal.z();
				texcoords[2 * v] = mv.tex_coords.x(); texcoords[2 * v + 1] = mv.tex_coords.y();
			}
			for (auto i : prim.indices) indices[ix++] = i;
			p++;
		}
	}
	for (size_t m = 0; m < sc.materials.size(); m++)
	{
		auto & mat = sc.materials[m];
		mat_bgra[4 * m] = mat.color.o.b; mat_bgra[4 * m + 1] = mat.color.o.g; mat_bgra[4 * m + 2] = mat.color.o.r; mat_bgra[4 * m + 3] = mat.color.o.a;
		mat_metal_rough[2 * m] = mat.metallic; mat_metal_rough[2 * m + 1] = mat.roughness;
		mat_tex_ds[2 * m] = mat.texture_idx; mat_tex_ds[2 * m + 1] = mat.double_sided ? 1 : 0;
	}
	for (size_t t = 0; t < sc.images.size(); t++)
	{
		tex_wh[2 * t] = (int)sc.images[t].m_mipmaps[0]->m_width;
		tex_wh[2 * t + 1] = (int)sc.images[t].m_mipmaps[0]->m_height;
	}
}

void ref_scene_export_texture(void * h, int idx, unsigned int * out_bgra)
{
	auto & mm = *static_cast<ref_scene *>(h)->scene.images[idx].m_mipmaps[0];
	std::memcpy(out_bgra, mm.m_bitmap, sizeof(unsigned int) * (size_t)mm.m_width * mm.m_height);
}

// scene_t::animations flattened: counts = animations, channels, steps
void ref_scene_animation_counts(void * h, unsigned * out3)
{
	auto & sc = static_cast<ref_scene *>(h)->scene;
	unsigned nc = 0, ns = 0;
	for (auto & a : sc.animations)
		for (auto & c : a.channels) { nc++; ns += (unsigned)c.steps.size(); }
	out3[0] = (unsigned)sc.animations.size(); out3[1] = nc; out3[2] = ns;
}

void ref_scene_export_animations(void * h, float * anim_end_time, int * chan_anim, int * chan_node, int * chan_path,
                                 unsigned * chan_first_step, unsigned * chan_n_steps, float * step_time, float * step_value4)
{
	auto & sc = static_cast<ref_scene *>(h)->scene;
	unsigned c = 0, st = 0;
	for (size_t a = 0; a < sc.animations.size(); a++)
	{
		anim_end_time[a] = sc.animations[a].end_time;
		for (auto & ch : sc.animations[a].channels)
		{
			chan_anim[c] = (int)a; chan_node[c] = ch.node_idx; chan_path[c] = (int)ch.path;
			chan_first_step[c] = st; chan_n_steps[c] = (unsigned)ch.steps.size();
			for (auto & step : ch.steps)
			{
				step_time[st] = step.time;
				step_value4[4 * st] = step.value.x(); step_value4[4 * st + 1] = step.value.y();
				step_value4[4 * st + 2] = step.value.z(); step_value4[4 * st + 3] = step.value.w();
				st++;
			}
			c++;
		}
	}
}

// the inverse, for scenes built with ref_scene_import (replaces any animations the scene had)
void ref_scene_import_animations(void * h, unsigned n_anims, const float * anim_end_time, unsigned n_channels, const int * chan_anim,
                                 const int * chan_node, const int * chan_path, const unsigned * chan_first_step,
                                 const unsigned * chan_n_steps, const float * step_time, const float * step_value4)
{
	auto & sc = static_cast<ref_scene *>(h)->scene;
	sc.animations.clear();
	sc.animations.resize(n_anims);
	for (unsigned a = 0; a < n_anims; a++) sc.animations[a].end_time = anim_end_time[a];
	for (unsigned c = 0; c < n_channels; c++)
	{
		auto & ch = sc.animations[chan_anim[c]].channels.emplace_back(
			swegl::animation_channel_t{chan_node[c], (swegl::animation_channel_t::path_t)chan_path[c], {}});
		for (unsigned k = chan_first_step[c]; k < chan_first_step[c] + chan_n_steps[c]; k++)
			ch.steps.emplace_back(swegl::animation_step_t{step_time[k],
				swegl::vec4f_t(step_value4[4 * k], step_value4[4 * k + 1], step_value4[4 * k + 2], step_value4[4 * k + 3])});
	}
}

// node TRS as scene_t::animate left them (rotation 4x4, scale 3, translation 3 per node)
void ref_scene_node_trs(void * h, float * node_scale, float * node_rotation, float * node_translation)
{
	auto & sc = static_cast<ref_scene *>(h)->scene;
	for (size_t n = 0; n < sc.nodes.size(); n++)
	{
		auto & node = sc.nodes[n];
		node_scale[3 * n] = node.scale.x(); node_scale[3 * n + 1] = node.scale.y(); node_scale[3 * n + 2] = node.scale.z();
		for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) node_rotation[16 * n + 4 * r + c] = node.rotation[r][c];
		node_translation[3 * n] = node.translation.x(); node_translation[3 * n + 1] = node.translation.y(); node_translation[3 * n + 2] = node.translation.z();
	}
}

void ref_scene_animate(void * h, float seconds) { static_cast<ref_scene *>(h)->scene.animate(seconds); }

// per-frame node state after render(): original_to_world_matrix (16) and scale(rotation, scale) 3x3 (9)
void ref_scene_node_matrices(void * h, float * node_world16, float * node_normal9)
{
	auto & sc = static_cast<ref_scene *>(h)->scene;
	for (size_t n = 0; n < sc.nodes.size(); n++)
	{
		auto & node = sc.nodes[n];
		for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) node_world16[16 * n + 4 * r + c] = node.original_to_world_matrix[r][c];
		swegl::matrix44_t rs = swegl::scale(node.rotation, node.scale);
		for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) node_normal9[9 * n + 3 * r + c] = rs[r][c];
	}
}

// post-render vertex state, flattened like ref_scene_export
void ref_scene_vertex_state(void * h, float * v_world, float * v_viewport, float * normal_world, unsigned char * yes)
{
	auto * s = static_cast<ref_scene *>(h);
	unsigned v = 0;
	for (auto & node : s->scene.nodes)
		for (auto & prim : node.primitives)
		{
			unsigned nv = (unsigned)prim.vertices.size() - s->spare;
			for (unsigned k = 0; k < nv; k++, v++)
			{
				auto & mv = prim.vertices[k];
				v_world[3 * v] = mv.v_world.x(); v_world[3 * v + 1] = mv.v_world.y(); v_world[3 * v + 2] = mv.v_world.z();
				v_viewport[3 * v] = mv.v_viewport.x(); v_viewport[3 * v + 1] = mv.v_viewport.y(); v_viewport[3 * v + 2] = mv.v_viewport.z();
				normal_world[3 * v] = mv.normal_world.x(); normal_world[3 * v + 1] = mv.normal_world.y(); normal_world[3 * v + 2] = mv.normal_world.z();
				yes[v] = mv.yes ? 1 : 0;
			}
		}
}

void * ref_screen_new(int w, int h)
{
	auto * sc = new ref_screen;
	sc->pixels.assign((size_t)w * h, 0u);
	sc->surface.w = w; sc->surface.h = h; sc->surface.pitch = w * 4;
	sc->surface.pixels = sc->pixels.data();
	sc->surface.format = &sc->format;
	return sc;
}
void ref_screen_free(void * h) { delete static_cast<ref_screen *>(h); }
unsigned int * ref_screen_pixels(void * h) { return static_cast<ref_screen *>(h)->pixels.data(); }

// light_mode 0 none / 1 flat / 2 phong, tex_mode 0 plain / 1 nearest / 2 bilinear
void * ref_viewport_new(void * screen, int x, int y, int w, int h, int light_mode, int tex_mode, int layers)
{
	using namespace swegl;
	auto * v = new ref_viewport;
	v->screen = static_cast<ref_screen *>(screen);
	if (light_mode == 1) v->shader = make_combined<pixel_shader_lights_flat>(tex_mode);
	else if (light_mode == 2) v->shader = make_combined<pixel_shader_lights_phong>(tex_mode);
	else if (tex_mode == 0) v->shader = std::make_shared<pixel_shader_t>();
	else if (tex_mode == 1) v->shader = std::make_shared<pixel_shader_texture>();
	else v->shader = std::make_shared<pixel_shader_texture_bilinear>();
	v->vp = std::make_unique<viewport_t>(x, y, w, h, &v->screen->surface, v->shader, layers);
	v->vp->set_post_shader(v->post_null); // src/test_1.cpp:356-357
	return v;
}
void ref_viewport_free(void * h) { delete static_cast<ref_viewport *>(h); }

// src/test_1.cpp:355: post_shader_depth_box(focal_distance, focal_depth, viewport).  In libswegl_ref.so this is the
// reference's DoF as shipped (reads out of bounds: do not call it there); libswegl_ref_dofr.so is built from the header
// repaired by oracle/dof_r.patch; in libswegl_dropin.so the adapter recognises the type and runs DoF-R on the device.
void ref_viewport_set_dof(void * h, float focal_distance, float focal_depth)
{
	auto * v = static_cast<ref_viewport *>(h);
	v->post_dof = std::make_unique<swegl::post_shader_depth_box>(focal_distance, focal_depth, *v->vp);
	v->vp->set_post_shader(*v->post_dof);
}

// op: 0 translate(a,b,c)  1 rotate_x(a)  2 rotate_y(a)  3 rotate_z(a)   (src/projection/camera.cpp:37-58)
void ref_viewport_camera(void * h, int op, float a, float b, float c)
{
	auto & cam = static_cast<ref_viewport *>(h)->vp->camera();
	switch (op)
	{
		case 0: cam.translate(a, b, c); break;
		case 1: cam.rotate_x(a); break;
		case 2: cam.rotate_y(a); break;
		case 3: cam.rotate_z(a); break;
	}
}

void ref_viewport_get(void * h, float * view16, float * proj16, float * cam_pos3, float * vpm4)
{
	auto & vp = *static_cast<ref_viewport *>(h)->vp;
	for (int r = 0; r < 4; r++)
		for (int c = 0; c < 4; c++)
		{
			view16[4 * r + c] = vp.camera().m_viewmatrix[r][c];
			proj16[4 * r + c] = vp.camera().m_projectionmatrix[r][c];
		}
	cam_pos3[0] = vp.camera().m_center.x(); cam_pos3[1] = vp.camera().m_center.y(); cam_pos3[2] = vp.camera().m_center.z();
	vpm4[0] = vp.m_viewportmatrix[0][0]; vpm4[1] = vp.m_viewportmatrix[0][3];
	vpm4[2] = vp.m_viewportmatrix[1][1]; vpm4[3] = vp.m_viewportmatrix[1][3];
}

float * ref_viewport_zbuffer(void * h) { return static_cast<ref_viewport *>(h)->vp->zbuffer(); }

void ref_render(void * scene, void * viewport)
{
	swegl::render(static_cast<ref_scene *>(scene)->scene, *static_cast<ref_viewport *>(viewport)->vp);
}

void ref_render2(void * scene, void * vp1, void * vp2)
{
	swegl::render(static_cast<ref_scene *>(scene)->scene, *static_cast<ref_viewport *>(vp1)->vp, *static_cast<ref_viewport *>(vp2)->vp);
}

void ref_render4(void * scene, void * vp1, void * vp2, void * vp3, void * vp4)
{
	swegl::render(static_cast<ref_scene *>(scene)->scene, *static_cast<ref_viewport *>(vp1)->vp, *static_cast<ref_viewport *>(vp2)->vp,
	              *static_cast<ref_viewport *>(vp3)->vp, *static_cast<ref_viewport *>(vp4)->vp);
}

// times `frames` calls of swegl::render after `warmup` untimed ones; returns total seconds and
// writes per-frame milliseconds when out_ms != null
double ref_time_render(void * scene, void * viewport, int warmup, int frames, double * out_ms)
{
	auto & sc = static_cast<ref_scene *>(scene)->scene;
	auto & vp = *static_cast<ref_viewport *>(viewport)->vp;
	for (int i = 0; i < warmup; i++) swegl::render(sc, vp);
	double total = 0;
	for (int i = 0; i < frames; i++)
	{
		auto t0 = std::chrono::steady_clock::now();
		swegl::render(sc, vp);
		auto t1 = std::chrono::steady_clock::now();
		double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
		if (out_ms) out_ms[i] = ms;
		total += ms / 1000.0;
	}
	return total;
}

#ifdef SWEGL_B200_DROPIN
static std::string g_dropin_error;
const char * ref_dropin_last_error() { return g_dropin_error.c_str(); }
// every entry point below returns 0, or 1 with the exception's text in ref_dropin_last_error() (a C++ exception must not
// cross the C boundary into ctypes)
#define DROPIN_GUARD(body) try { body; g_dropin_error.clear(); return 0; } catch (const std::exception & e) { g_dropin_error = e.what(); return 1; }
// Only in libswegl_dropin.so: the multi-context C++ hosts of swegl_b200/host/swegl_b200_host.hpp driven with the
// reference's own scene_t / viewport_t objects (tests/test_dropin_gpu.py).
// `frames` frames of one viewport through swegl_b200::pipeline_t (depth contexts, round robin); the camera turns by
// `dyaw` before every frame; the LAST frame is collected into the viewport's surface
int ref_render_pipelined(void * scene, void * viewport, int depth, int frames, float dyaw)
{
	auto & sc = static_cast<ref_scene *>(scene)->scene;
	auto & vp = *static_cast<ref_viewport *>(viewport)->vp;
	DROPIN_GUARD(
		swegl_b200::pipeline_t pipe(0, depth);
		int slot = 0;
		for (int i = 0; i < frames; i++)
		{
			vp.camera().rotate_y(dyaw);
			slot = pipe.submit(sc, vp);
		}
		pipe.collect(slot, vp);
		pipe.synchronize())
}

// one frame of one viewport in `n_ctx` row bands (contexts of device 0 standing in for GPUs), `frames` times
int ref_render_sharded(void * scene, void * viewport, int n_ctx, int frames)
{
	auto & sc = static_cast<ref_scene *>(scene)->scene;
	auto & vp = *static_cast<ref_viewport *>(viewport)->vp;
	DROPIN_GUARD(
		swegl_b200::sharded_renderer_t sh(std::vector<int>((size_t)n_ctx, 0));
		for (int i = 0; i < frames; i++) sh.render(sc, vp))
}

// the application's pair  scene.animate(t); swegl::render(scene, viewport)  (src/test_1.cpp:374-378) with the animation and
// the node-hierarchy product evaluated on the device (swegl_b200::render_animated): the host scene_t is not touched
int ref_render_animated(void * scene, void * viewport, float elapsed_seconds)
{
	DROPIN_GUARD(swegl_b200::render_animated(static_cast<ref_scene *>(scene)->scene, elapsed_seconds, *static_cast<ref_viewport *>(viewport)->vp))
}

// swegl::render(scene, vp1, vp2, vp3, vp4) with one viewport per context
int ref_render_sharded4(void * scene, void * vp1, void * vp2, void * vp3, void * vp4, int n_ctx)
{
	auto & sc = static_cast<ref_scene *>(scene)->scene;
	DROPIN_GUARD(
		swegl_b200::sharded_renderer_t sh(std::vector<int>((size_t)n_ctx, 0));
		sh.render(sc, *static_cast<ref_viewport *>(vp1)->vp, *static_cast<ref_viewport *>(vp2)->vp,
		          *static_cast<ref_viewport *>(vp3)->vp, *static_cast<ref_viewport *>(vp4)->vp))
}
#endif

} // extern "C"
