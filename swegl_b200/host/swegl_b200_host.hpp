// swegl_b200_host.hpp — the fast paths of the C++ host side: more than one blocking context on device 0.
//
// swegl's frame loop (src/test_1.cpp:366-385) is `swegl::render(scene, viewports...)` once per frame, and
// swegl::render itself is "original_to_world once, then _render per viewport" (swegl/render/renderer.hpp:18-34).
// The drop-in of swegl_b200_adapter.hpp maps that 1:1 onto one context.  This header adds, on the same swegl objects
// (scene_t, viewport_t; compiled against the user's swegl checkout) and over the same C ABI:
//
//   pipeline_t           N contexts on one GPU, independent frames round robin: the latency-bound head of frame i+1
//                        (vertex, mark, set-up, spans) runs under the fragment / DoF kernels of frame i.
//   sharded_renderer_t   ONE process, one context per GPU (NVLink peer access):
//                          render(scene, vp)            one frame in sort-first row bands: every GPU
//                                                       draws its band of the FULL viewport and its last kernel stores
//                                                       the finished rows into GPU 0's screen; flags in GPU 0's memory
//                                                       instead of a collective (swegl_b200_set_frame_sync);
//                          render(scene, vp1, vp2, ...)  renderer.hpp:20-34's viewport loop with viewport v on GPU
//                                                       v mod N (BASELINE.json config 4).
//                        The frame lands where swegl::render leaves it: vp.m_screen->pixels and vp.zbuffer().
//
// Errors: std::runtime_error, like the adapter.  No CPU fallback.
#pragma once

#include <algorithm>
#include <memory>
#include <vector>

#include "swegl_b200_adapter.hpp"

namespace swegl_b200
{

// ---------------------------------------------------------------------------------------------------------------
class pipeline_t
{
public:
	pipeline_t(int device, int depth)
	{
		for (int i = 0; i < std::max(1, depth); i++) m_engines.push_back(std::make_unique<engine_t>(device));
		if (m_engines.size() > 1)           // the contexts' frames overlap on the GPU: kernels that hold fewer SM resources
			for (auto & e : m_engines) e->check(swegl_b200_set_shared_gpu(e->ctx(), 1), "set_shared_gpu");
	}

	int depth() const { return (int)m_engines.size(); }

	// queue one frame (all its viewports) on the next context and return that context's index.  Nothing is read
	// back: the frame stays in that context's device screen until the context is used again, depth() submits later.
	template <typename... T>
	int submit(swegl::scene_t & scene, T &... viewports)
	{
		const int k = m_next;
		engine_t & e = *m_engines[k];
		e.begin_frame(scene, false);
		(queue(e, viewports), ...);
		m_next = (k + 1) % depth();
		return k;
	}

	// wait for context `slot` and copy its frame (the viewports' rectangles) into their surfaces
	template <typename... T>
	void collect(int slot, T &... viewports)
	{
		engine_t & e = *m_engines[slot];
		e.check(swegl_b200_synchronize(e.ctx()), "synchronize");
		(read(e, viewports), ...);
	}

	void synchronize()
	{
		for (auto & e : m_engines) e->check(swegl_b200_synchronize(e->ctx()), "synchronize");
	}

	engine_t & engine(int slot) { return *m_engines[slot]; }

private:
	std::vector<std::unique_ptr<engine_t>> m_engines;
	int m_next = 0;

	static void queue(engine_t & e, swegl::viewport_t & vp)
	{
		e.ensure_screen(vp.m_screen);
		e.render_viewport_device(engine_t::describe(vp));
	}
	static void read(engine_t & e, swegl::viewport_t & vp)
	{
		e.check(swegl_b200_read_rect(e.ctx(), vp.m_x, vp.m_y, vp.m_w, vp.m_h, vp.m_screen->pixels, vp.m_screen->pitch), "read_rect");
		// (the depth of the LAST viewport rendered by the context is what it still holds)
	}
};

// ---------------------------------------------------------------------------------------------------------------
class sharded_renderer_t
{
public:
	// one context per entry of `devices` (an entry may repeat: several contexts of one GPU behave like several GPUs,
	// which is how the protocol is tested on a one-GPU box)
	explicit sharded_renderer_t(const std::vector<int> & devices)
	{
		if (devices.empty()) throw std::runtime_error("swegl_b200: sharded_renderer_t needs at least one device");
		for (int d : devices) m_engines.push_back(std::make_unique<engine_t>(d));
		for (auto & e : m_engines)
			e->check(swegl_b200_enable_peer(e->ctx(), devices[0]), "enable_peer");      // everybody stores into GPU 0's screen
	}

	int world() const { return (int)m_engines.size(); }

	// ONE viewport, row bands: band r of the full viewport on context r
	void render(swegl::scene_t & scene, swegl::viewport_t & vp)
	{
		const int n = world();
		prepare(scene, vp.m_screen);
		swegl_b200_viewport_desc base = engine_t::describe(vp);
		if (base.transparency_layers > 0 && n > 1 && swegl_b200_scene_opaque(m_engines[0]->ctx()) != 1)     // (opaque scenes: the layers are the identity)
			throw std::runtime_error("swegl_b200: transparency layers are not sharded (render the viewport on one context)");
		if (m_bands.size() != (size_t)n + 1 || m_bands_h != vp.m_h) even_bands(vp.m_h);
		std::vector<swegl_b200_viewport_desc> d(n, base);
		for (int r = 0; r < n; r++) { d[r].band_y0 = m_bands[r]; d[r].band_y1 = m_bands[r + 1]; }
		if (n == 1) { d[0].band_y0 = d[0].band_y1 = 0; }

		// the first frame of a configuration runs synchronously on every context: it sizes the span / piece / fragment
		// pools (a frame of the protocol cannot be redone on one rank alone)
		const bool sized = m_sized_w == vp.m_w && m_sized_h == vp.m_h && m_sized_post == base.post_mode;
		if (!sized)
		{
			disarm();
			for (int r = 0; r < n; r++) { swegl_b200_stats st; m_engines[r]->render_viewport_device(d[r], &st); }
			m_sized_w = vp.m_w; m_sized_h = vp.m_h; m_sized_post = base.post_mode;
		}
		arm();
		for (int attempt = 0; attempt < 3; attempt++)
		{
			if (attempt) for (int r = 0; r < n; r++) m_engines[r]->begin_frame(scene, true);
			for (int r = 0; r < n; r++) m_engines[r]->render_viewport_device(d[r]);
			int worst = SWEGL_B200_OK;
			for (int r = 0; r < n; r++)
			{
				const int rc = swegl_b200_synchronize(m_engines[r]->ctx());         // context 0 passing its stream's end = the frame is assembled
				if (rc == SWEGL_B200_ERR_CAPACITY) worst = rc;
				else if (rc) m_engines[r]->check(rc, "synchronize");
			}
			if (worst == SWEGL_B200_OK) break;
			// a pool overflowed somewhere (now enlarged): restart the protocol's frame counter on every rank and redo
			disarm(); arm();
			if (attempt == 2) throw std::runtime_error("swegl_b200: the span / piece pools keep overflowing");
		}
		engine_t & e0 = *m_engines[0];
		e0.check(swegl_b200_read_rect(e0.ctx(), vp.m_x, vp.m_y, vp.m_w, vp.m_h, vp.m_screen->pixels, vp.m_screen->pitch), "read_rect");
		if (float * z = vp.zbuffer())
			for (int r = 0; r < n; r++)
				m_engines[r]->check(swegl_b200_read_depth_rows(m_engines[r]->ctx(), m_bands[r], m_bands[r + 1], z), "read_depth_rows");
	}

	// several viewports of one surface (renderer.hpp:20-34): viewport v on context v mod N, every context storing its
	// finished pixels into context 0's screen
	template <typename... T>
	void render(swegl::scene_t & scene, swegl::viewport_t & first, swegl::viewport_t & second, T &... rest)
	{
		swegl::viewport_t * vps[] = { &first, &second, &rest... };
		const int nv = (int)(sizeof(vps) / sizeof(vps[0])), n = world();
		prepare(scene, first.m_screen);
		disarm();
		target_rank0();
		// rounds of up to N viewports, one per context, queued asynchronously so that the GPUs work side by side; a context
		// keeps the depth of its last viewport only, so depth is read at the end of every round
		for (int v0 = 0; v0 < nv; v0 += n)
		{
			const int m = std::min(n, nv - v0);
			for (int r = 0; r < m; r++) m_engines[r]->render_viewport_device(engine_t::describe(*vps[v0 + r]));
			for (int r = 0; r < m; r++)
			{
				engine_t & e = *m_engines[r];
				const int rc = swegl_b200_synchronize(e.ctx());
				if (rc == SWEGL_B200_ERR_CAPACITY)
				{
					swegl_b200_stats st;                                               // pools were too small (now enlarged): redo, synchronously
					e.begin_frame(scene, true);
					e.render_viewport_device(engine_t::describe(*vps[v0 + r]), &st);
				}
				else if (rc) e.check(rc, "synchronize");
				if (float * z = vps[v0 + r]->zbuffer())
					e.check(swegl_b200_read_depth_rows(e.ctx(), 0, vps[v0 + r]->m_h, z), "read_depth_rows");
			}
		}
		engine_t & e0 = *m_engines[0];
		for (int v = 0; v < nv; v++)
			e0.check(swegl_b200_read_rect(e0.ctx(), vps[v]->m_x, vps[v]->m_y, vps[v]->m_w, vps[v]->m_h, vps[v]->m_screen->pixels, vps[v]->m_screen->pitch), "read_rect");
	}

	// row cuts of the next banded frames (n + 1 ascending values from 0 to the viewport height), e.g. from the ranks'
	// measured times; the default is an even split
	void set_bands(const std::vector<int> & cuts, int viewport_h) { m_bands = cuts; m_bands_h = viewport_h; m_sized_w = -1; }

	engine_t & engine(int r) { return *m_engines[r]; }

private:
	std::vector<std::unique_ptr<engine_t>> m_engines;
	std::vector<int> m_bands;
	int m_bands_h = -1, m_sized_w = -1, m_sized_h = -1, m_sized_post = -1;
	bool m_armed = false, m_targeted = false;

	void prepare(swegl::scene_t & scene, const SDL_Surface * screen)
	{
		bool first = true;
		for (auto & e : m_engines)
		{
			const int w0 = screen->w, h0 = screen->h;
			void * before = nullptr; swegl_b200_device_buffers(e->ctx(), &before, nullptr);
			e->ensure_screen(screen);
			void * after = nullptr; swegl_b200_device_buffers(e->ctx(), &after, nullptr);
			if (before != after) { m_armed = false; m_targeted = false; m_sized_w = -1; (void)w0; (void)h0; }      // a new screen: targets and flags are stale
			e->begin_frame(scene, !first);                                             // the hierarchy product runs once (vertex_shaders.hpp:16-18)
			first = false;
		}
	}
	void even_bands(int h)
	{
		const int n = world();
		m_bands.assign(n + 1, 0);
		for (int r = 0; r <= n; r++) m_bands[r] = (int)((long long)h * r / n);
		m_bands_h = h;
	}
	void target_rank0()
	{
		if (m_targeted) return;
		void * screen0 = nullptr;
		m_engines[0]->check(swegl_b200_device_buffers(m_engines[0]->ctx(), &screen0, nullptr), "device_buffers");
		for (int r = 1; r < world(); r++) m_engines[r]->check(swegl_b200_set_color_target(m_engines[r]->ctx(), screen0), "set_color_target");
		m_targeted = true;
	}
	void arm()
	{
		if (m_armed || world() == 1) return;
		target_rank0();
		for (int r = 0; r < world(); r++) m_engines[r]->check(swegl_b200_set_frame_sync(m_engines[r]->ctx(), r, world()), "set_frame_sync");
		m_armed = true;
	}
	void disarm()
	{
		if (!m_armed) return;
		for (auto & e : m_engines) e->check(swegl_b200_set_frame_sync(e->ctx(), -1, 0), "set_frame_sync");
		m_armed = false;
	}
};

} // namespace swegl_b200
