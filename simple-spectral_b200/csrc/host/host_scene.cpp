// host_scene.cpp — the reference's hard-coded scenes and Scene::_init (reference src/scene.cpp:16-415), built
// directly as the flat quad list of the C ABI.  Geometry literals are scene DATA of the Cornell box
// (http://www.graphics.cornell.edu/online/box/data.html, as amended by the reference: ceiling cut around the
// light, scene.cpp:130-193); they are tabulated here, not transcribed statement by statement.
#include <cmath>
#include <cstring>

#include "host_math.hpp"
#include "ssb_host.hpp"

namespace ssbh {

bool Material::is_emissive() const { return Spectrum::integrate(emission) > 0.0f; }  // material.cpp:100-106

namespace {

struct V { float x, y, z, s, t; };

ssb_vertex vert(V const& v) { return ssb_vertex{ { v.x, v.y, v.z }, { v.s, v.t } }; }
ssb_tri tri(V const& a, V const& b, V const& c) {  // PrimTri ctor, geometry.hpp:60-69
	ssb_tri t{};
	t.v[0] = vert(a); t.v[1] = vert(b); t.v[2] = vert(c);
	F3 n = normalize(cross(f3(b.x, b.y, b.z) - f3(a.x, a.y, a.z), f3(c.x, c.y, c.z) - f3(a.x, a.y, a.z)));
	t.normal[0] = n.x; t.normal[1] = n.y; t.normal[2] = n.z;
	return t;
}

class Builder {
public:
	Scene scene;
	ColorData const& color;
	explicit Builder(ColorData const& c) : color(c) {}

	Spectrum constant(float v) const { return Spectrum(v, color.lambda_min, color.lambda_max); }
	int add_material(std::string const& name, Material m) {
		m.name = name;
		if (m.emission.data.empty()) m.emission = constant(0.0f);  // MaterialBase(): emission(0.0f)
		if (m.albedo_mode == SSB_ALBEDO_CONSTANT && m.albedo.data.empty()) m.albedo = constant(1.0f);  // Albedo()
		pending_.push_back(m);
		return static_cast<int>(pending_.size()) - 1;
	}
	// PrimQuad(material, v00, v10, v11, v01) = tri0(v00,v10,v11) + tri1(v00,v11,v01), geometry.hpp:86-96
	void quad(int material, V const& v00, V const& v10, V const& v11, V const& v01) {
		ssb_quad q{};
		q.tri[0] = tri(v00, v10, v11);
		q.tri[1] = tri(v00, v11, v01);
		q.material = static_cast<uint32_t>(material);
		quads_.push_back(q);
	}
	Material& pending(int i) { return pending_[static_cast<size_t>(i)]; }
	std::vector<ssb_quad>& quads() { return quads_; }

	// materials are emitted in first-use order over the primitive list; is_light is taken at finish time
	// (PrimBase's ctor evaluates material->is_emissive() when the primitive is created, geometry.cpp:7-9; no
	// scene changes emissiveness afterwards)
	Scene finish() {
		std::vector<int> remap(pending_.size(), -1);
		for (ssb_quad& q : quads_) {
			int m = static_cast<int>(q.material);
			if (remap[static_cast<size_t>(m)] < 0) {
				remap[static_cast<size_t>(m)] = static_cast<int>(scene.materials.size());
				scene.materials.push_back(pending_[static_cast<size_t>(m)]);
			}
			q.material = static_cast<uint32_t>(remap[static_cast<size_t>(m)]);
			q.is_light = scene.materials[q.material].is_emissive() ? 1u : 0u;
		}
		scene.quads = quads_;
		for (uint32_t i = 0; i < scene.quads.size(); ++i)
			if (scene.quads[i].is_light) scene.lights.push_back(i);  // Scene::_init, scene.cpp:26-29
		if (scene.lights.empty()) throw Error{ -1, "scene has no lights" };
		scene.flatten();
		return scene;
	}

private:
	std::vector<Material> pending_;
	std::vector<ssb_quad> quads_;
};

void init_camera(Camera& cam) {  // Scene::_init, scene.cpp:16-24
	M4<float> P = perspectiveFov(cam.vfov_deg * static_cast<float>(0.01745329251994329576923690768489),
	                             static_cast<float>(cam.res[0]), static_cast<float>(cam.res[1]), cam.near_, cam.far_);
	F3 pos = f3(cam.pos[0], cam.pos[1], cam.pos[2]), dir = f3(cam.dir[0], cam.dir[1], cam.dir[2]), up = f3(cam.up[0], cam.up[1], cam.up[2]);
	M4<float> Vw = lookAt(pos, pos + dir, up);
	M4<double> Pd = widen(P), Vd = widen(Vw);
	M4<double> PVi = inverse(mul(Pd, Vd));
	std::memcpy(cam.matr_P, Pd.m, sizeof(Pd.m));
	std::memcpy(cam.matr_V, Vd.m, sizeof(Vd.m));
	std::memcpy(cam.matr_PV_inv, PVi.m, sizeof(PVi.m));
}

struct CornellIds { int floorceil, light, back, green, red, blocks; };

// Scene::get_new_cornell, scene.cpp:32-287
CornellIds build_cornell(Builder& b, std::string const& data_root) {
	Camera& cam = b.scene.camera;
	cam.pos[0] = 278; cam.pos[1] = 273; cam.pos[2] = -800;
	F3 d = normalize(f3(0, 0, 1));
	cam.dir[0] = d.x; cam.dir[1] = d.y; cam.dir[2] = d.z;
	cam.up[0] = 0; cam.up[1] = 1; cam.up[2] = 0;
	cam.res[0] = 512; cam.res[1] = 512; cam.near_ = 0.1f; cam.far_ = 1.0f; cam.vfov_deg = 39.0f;

	auto wgr = load_spectral_data(data_root + "/data/scenes/cornell/white-green-red.csv");
	if (wgr.size() != 3) throw Error{ -1, "Invalid data in file!" };
	auto lcsv = load_spectral_data(data_root + "/data/scenes/cornell/light.csv");
	if (lcsv.size() != 1) throw Error{ -1, "Invalid data in file!" };
	CornellIds id{};
	Material white; white.albedo = Spectrum(wgr[0], 400, 700);
	id.back = b.add_material("white-back", white);
	id.blocks = b.add_material("white-blocks", white);
	id.floorceil = b.add_material("white-floorceil", white);
	Material green; green.albedo = Spectrum(wgr[1], 400, 700);
	green.albedo_rgb[0] = 0.07f; green.albedo_rgb[1] = 0.38f; green.albedo_rgb[2] = 0.07f;  // scene.cpp:73 ("set heuristically")
	id.green = b.add_material("green", green);
	Material red; red.albedo = Spectrum(wgr[2], 400, 700);
	red.albedo_rgb[0] = 1; red.albedo_rgb[1] = 0; red.albedo_rgb[2] = 0;  // scene.cpp:76
	id.red = b.add_material("red", red);
	Material light;
	light.emission = Spectrum(lcsv[0], 400, 700) * 200.0f;
	light.albedo = b.constant(0.78f);
	for (int k = 0; k < 3; ++k) { light.emission_rgb[k] = 1.0f * 200.0f; light.albedo_rgb[k] = 0.78f; }  // scene.cpp:97-98
	id.light = b.add_material("light", light);

	const float Y = 548.8f;
	// floor
	b.quad(id.floorceil, { 552.8f, 0, 0, 1, 0 }, { 0, 0, 0, 0, 0 }, { 0, 0, 559.2f, 0, 1 }, { 549.6f, 0, 559.2f, 1, 1 });
	// ceiling corners A..D and light corners E..H (scene.cpp:141-148)
	V A{ 0, Y, 559.2f, 0, 0 }, B{ 556, Y, 559.2f, 0, 0 }, C{ 0, Y, 0, 0, 0 }, D{ 556, Y, 0, 0, 0 };
	V E{ 213, Y, 332, 0, 0 }, F{ 343, Y, 332, 0, 0 }, G{ 213, Y, 227, 0, 0 }, H{ 343, Y, 227, 0, 0 };
	auto st = [](V v, float s, float t) { v.s = s; v.t = t; return v; };
	b.quad(id.light, st(H, 1, 0), st(F, 1, 1), st(E, 0, 1), st(G, 0, 0));
	b.quad(id.floorceil, D, B, F, H);
	b.quad(id.floorceil, B, A, E, F);
	b.quad(id.floorceil, A, C, G, E);
	b.quad(id.floorceil, C, D, H, G);
	// back, right (green), left (red) walls
	b.quad(id.back, { 549.6f, 0, 559.2f, 0, 0 }, { 0, 0, 559.2f, 1, 0 }, { 0, Y, 559.2f, 1, 1 }, { 556, Y, 559.2f, 0, 1 });
	b.quad(id.green, { 0, 0, 559.2f, 1, 0 }, { 0, 0, 0, 0, 0 }, { 0, Y, 0, 0, 1 }, { 0, Y, 559.2f, 1, 1 });
	b.quad(id.red, { 552.8f, 0, 0, 0, 0 }, { 549.6f, 0, 559.2f, 1, 0 }, { 556, Y, 559.2f, 1, 1 }, { 556, Y, 0, 0, 1 });
	// blocks: a top quad, then four sides, each side (p_bottom, p_top, q_top, q_bottom)
	auto block = [&](float h, float const top[4][2], int const sides[4][2]) {
		b.quad(id.blocks, { top[0][0], h, top[0][1], 0, 0 }, { top[1][0], h, top[1][1], 0, 0 }, { top[2][0], h, top[2][1], 0, 0 }, { top[3][0], h, top[3][1], 0, 0 });
		for (int s = 0; s < 4; ++s) {
			float const* p = top[sides[s][0]];
			float const* q = top[sides[s][1]];
			b.quad(id.blocks, { p[0], 0, p[1], 0, 0 }, { p[0], h, p[1], 0, 0 }, { q[0], h, q[1], 0, 0 }, { q[0], 0, q[1], 0, 0 });
		}
	};
	float const short_top[4][2] = { { 130, 65 }, { 82, 225 }, { 240, 272 }, { 290, 114 } };
	int const short_sides[4][2] = { { 3, 2 }, { 0, 3 }, { 1, 0 }, { 2, 1 } };
	block(165.0f, short_top, short_sides);
	float const tall_top[4][2] = { { 423, 247 }, { 265, 296 }, { 314, 456 }, { 472, 406 } };
	int const tall_sides[4][2] = { { 0, 3 }, { 3, 2 }, { 2, 1 }, { 1, 0 } };
	block(330.0f, tall_top, tall_sides);
	return id;
}

}  // namespace

void Scene::flatten() {
	flat_materials.resize(materials.size());
	for (size_t m = 0; m < materials.size(); ++m) {
		ssb_material& f = flat_materials[m];
		f = ssb_material{};
		f.kind = materials[m].kind;
		f.albedo_mode = materials[m].albedo_mode;
		if (materials[m].albedo_mode == SSB_ALBEDO_CONSTANT) f.albedo = materials[m].albedo.flat();
		f.texture = materials[m].texture < 0 ? 0u : static_cast<uint32_t>(materials[m].texture);
		f.emission = materials[m].emission.flat();
		for (int k = 0; k < 3; ++k) { f.albedo_rgb[k] = materials[m].albedo_rgb[k]; f.emission_rgb[k] = materials[m].emission_rgb[k]; }
	}
	flat_textures.resize(textures.size());
	for (size_t t = 0; t < textures.size(); ++t) flat_textures[t] = ssb_texture{ textures[t].rgb8.data(), textures[t].width, textures[t].height };
	flat = ssb_scene{};
	std::memcpy(flat.camera.pv_inv, camera.matr_PV_inv, sizeof(flat.camera.pv_inv));
	std::memcpy(flat.camera.pos, camera.pos, sizeof(flat.camera.pos));
	std::memcpy(flat.camera.dir, camera.dir, sizeof(flat.camera.dir));
	flat.quads = quads.data(); flat.nquads = static_cast<uint32_t>(quads.size());
	flat.materials = flat_materials.data(); flat.nmaterials = static_cast<uint32_t>(flat_materials.size());
	flat.textures = flat_textures.empty() ? nullptr : flat_textures.data();
	flat.ntextures = static_cast<uint32_t>(flat_textures.size());
}

Scene scene_new(std::string const& name, std::string const& data_root, ColorData const& color, bool explicit_light_sampling) {
	Builder b(color);
	b.scene.name = name;
	std::string const lizard = data_root + "/data/scenes/crystal-lizard-4096.png";
	if (name == "cornell") {
		build_cornell(b, data_root);
	} else if (name == "cornell-srgb") {  // Scene::get_new_cornell_srgb, scene.cpp:288-319
		CornellIds id = build_cornell(b, data_root);
		b.scene.textures.push_back(load_png_rgb8(lizard));
		Material tex; tex.albedo_mode = SSB_ALBEDO_TEXTURE; tex.texture = 0;
		int mtl_tex = b.add_material("srgb", tex);
		Material white1; white1.albedo = b.constant(1.0f);
		int mtl_white1 = b.add_material("white1", white1);
		for (ssb_quad& q : b.quads()) {
			int m = static_cast<int>(q.material);
			if (m == id.blocks || m == id.floorceil) q.material = static_cast<uint32_t>(mtl_white1);
			else if (m == id.red) q.material = static_cast<uint32_t>(mtl_tex);
		}
		b.pending(id.light).emission = color.D65_rad * 30.0f;
		for (int k = 0; k < 3; ++k) b.pending(id.light).emission_rgb[k] = 1.0f * 30.0f;  // scene.cpp:313-316
	} else if (name == "plane-srgb") {  // Scene::get_new_plane_srgb, scene.cpp:320-415
		Camera& cam = b.scene.camera;
		cam.pos[0] = 0; cam.pos[1] = 0; cam.pos[2] = 5;
		F3 d = normalize(f3(0, 0, 0) - f3(cam.pos[0], cam.pos[1], cam.pos[2]));
		cam.dir[0] = d.x; cam.dir[1] = d.y; cam.dir[2] = d.z;
		cam.up[0] = 0; cam.up[1] = 1; cam.up[2] = 0;
		cam.res[0] = 512; cam.res[1] = 512; cam.near_ = 0.1f; cam.far_ = 1.0f;
		cam.vfov_deg = (2.0f * std::atan2(1.0f, cam.pos[2])) * static_cast<float>(57.295779513082320876798154814105);
		Material light; light.albedo = b.constant(0.0f); light.emission = color.D65_rad;
		for (int k = 0; k < 3; ++k) { light.albedo_rgb[k] = 0.0f; light.emission_rgb[k] = 1.0f; }  // scene.cpp:340-341
		int mtl_light = b.add_material("light", light);
		b.scene.textures.push_back(load_png_rgb8(lizard));
		Material tex; tex.albedo_mode = SSB_ALBEDO_TEXTURE; tex.texture = 0;
		tex.kind = explicit_light_sampling ? SSB_MATERIAL_LAMBERT : SSB_MATERIAL_MIRROR;  // scene.cpp:346-355
		int mtl_tex = b.add_material("tex", tex);
		b.quad(mtl_tex, { -1, -1, 0, 0, 0 }, { 1, -1, 0, 1, 0 }, { 1, 1, 0, 1, 1 }, { -1, 1, 0, 0, 1 });
		const float s = 10.0f;
		float const box[6][4][3] = {
			{ { -s, -s, s }, { -s, -s, -s }, { -s, s, -s }, { -s, s, s } },
			{ { s, -s, -s }, { s, -s, s }, { s, s, s }, { s, s, -s } },
			{ { -s, -s, s }, { s, -s, s }, { s, -s, -s }, { -s, -s, -s } },
			{ { s, s, s }, { -s, s, s }, { -s, s, -s }, { s, s, -s } },
			{ { -s, -s, -s }, { s, -s, -s }, { s, s, -s }, { -s, s, -s } },
			{ { s, -s, s }, { -s, -s, s }, { -s, s, s }, { s, s, s } },
		};
		for (auto const& f : box)
			b.quad(mtl_light, { f[0][0], f[0][1], f[0][2], 0, 0 }, { f[1][0], f[1][1], f[1][2], 0, 0 }, { f[2][0], f[2][1], f[2][2], 0, 0 }, { f[3][0], f[3][1], f[3][2], 0, 0 });
	} else {
		throw Error{ -3, "Unrecognized scene \"" + name + "\"!  (Supported scenes: \"cornell\", \"cornell-srgb\", \"plane-srgb\")" };  // renderer.cpp:32-37
	}
	init_camera(b.scene.camera);
	return b.finish();
}

}  // namespace ssbh
