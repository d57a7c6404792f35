// Scene.cpp -- see Scene.h.  A from-scratch tokenising parser; the accepted grammar and its quirks follow
// src/Scene.cpp:62-410 of the reference (cited inline) because they decide the bytes of vert_data/mat_data.
#include "Scene.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <limits>
#include <map>
#include <stdexcept>

namespace yune
{
    namespace
    {
        // Whitespace-separated token cursor over one line (what `stringstream >> x` does in the reference).
        struct Cursor
        {
            const char* p; const char* end;
            explicit Cursor(const std::string& line) : p(line.data()), end(line.data() + line.size()) {}
            static bool space(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == '\v' || c == '\f'; }
            bool word(std::string& out)
            {
                while (p < end && space(*p)) p++;
                const char* b = p;
                while (p < end && !space(*p)) p++;
                out.assign(b, p);
                return p != b;
            }
            float real()        // failed extraction leaves 0, like num_get
            {
                std::string w; if (!word(w)) return 0.0f;
                char* stop = nullptr; float v = std::strtof(w.c_str(), &stop);
                return stop == w.c_str() ? 0.0f : v;
            }
            int integer()
            {
                std::string w; if (!word(w)) return 0;
                return (int)std::strtol(w.c_str(), nullptr, 10);
            }
        };

        bool skippable(const std::string& line) { return line.empty() || line[0] == '#'; }   // src/Scene.cpp:78, 179, 256

        // One MTL statement applied to the last material (src/Scene.cpp:185-225).  Note px/py are crossed, as in the reference.
        void applyMaterialKey(const std::string& key, Cursor& cur, Material& m)
        {
            auto colour = [&](Float4& dst) { float x = cur.real(), y = cur.real(), z = cur.real(); dst = {{x, y, z, 1.0f}}; };
            if (key == "ke") colour(m.ke);
            else if (key == "kd") colour(m.kd);
            else if (key == "ks") colour(m.ks);
            else if (key == "n") m.n = cur.real();
            else if (key == "k") m.k = cur.real();
            else if (key == "px") m.py = cur.real();          // sic (src/Scene.cpp:212-215)
            else if (key == "py") m.px = cur.real();
            else if (key == "alpha_x") m.alpha_x = cur.real();
            else if (key == "alpha_y") m.alpha_y = cur.real();
            else if (key == "is_specular") m.is_specular = cur.integer();
            else if (key == "is_transmissive") m.is_transmissive = cur.integer();
        }

        void parseMaterials(std::istream& file, std::vector<Material>& mats, std::map<std::string, int>* index)
        {
            std::string line, key;
            while (std::getline(file, line)) {
                if (skippable(line)) continue;
                Cursor cur(line);
                if (!cur.word(key)) continue;
                if (key == "newmtl") {
                    std::string id; cur.word(id);
                    if (index) (*index)[id] = (int)mats.size();
                    mats.push_back(newMaterial());
                } else if (!mats.empty())
                    applyMaterialKey(key, cur, mats.back());
            }
            if (mats.empty()) throw std::runtime_error("Bad Material file.");
        }

        // The .mtl name is only honoured when 'mtllib' is the first statement of the .obj (src/Scene.cpp:385-410).
        std::string materialFileOf(const std::string& filepath)
        {
            std::ifstream file(filepath);
            std::string line, key;
            while (file.is_open() && std::getline(file, line)) {
                if (skippable(line)) continue;
                Cursor cur(line);
                cur.word(key);
                if (key != "mtllib") return "";
                cur.word(key);
                return key;
            }
            return "";
        }

        struct P3 { float x, y, z; };
    }

    Scene::Scene() { clearValues(); }

    void Scene::clearValues()
    {
        num_triangles = 0; scene_size_kb = scene_size_mb = 0;
        mat_filename.clear(); mat_file.clear(); scene_file.clear();
        vert_data.clear(); mat_data.clear(); cpu_tri_list.clear();
    }

    void Scene::reloadMatFile()
    {
        std::ifstream file(mat_file);
        if (!file.is_open()) throw std::runtime_error("Error opening material file.");
        mat_data.clear();
        parseMaterials(file, mat_data, nullptr);
    }

    void Scene::loadModel(std::string filepath, std::string filename)
    {
        clearValues();
        std::map<std::string, int> mat_index;
        std::string mat_fn = materialFileOf(filepath);
        std::string mat_fp = filepath;
        mat_fp.erase(mat_fp.find_last_of("/") + 1);
        if (mat_fn.empty()) {
            // no mtllib: the reference creates <name>.mtl on disk with default values; we keep the default in memory
            mat_fn = filename;
            if (mat_fn.size() >= 3) mat_fn.replace(mat_fn.size() - 3, 3, "mtl");
            mat_fp += mat_fn;
            mat_index["default"] = 0;
            mat_data.push_back(newMaterial());
        } else {
            mat_fp += mat_fn;
            std::ifstream mfile(mat_fp);
            if (!mfile.is_open()) throw std::runtime_error("Error opening material file.");
            parseMaterials(mfile, mat_data, &mat_index);
        }

        std::ifstream file(filepath);
        if (!file.is_open()) throw std::runtime_error("Error opening object file...");

        const float inf = std::numeric_limits<float>::max();
        root.p_min = {{inf, inf, inf, 1.0f}};
        root.p_max = {{-inf, -inf, -inf, 1.0f}};
        std::vector<P3> vertices, normals;
        std::string line, key, current_mtl;
        int mat_id = -1;
        auto lookup = [&](const std::string& name) { auto it = mat_index.find(name); return it == mat_index.end() ? 0 : it->second; };

        while (std::getline(file, line)) {
            if (skippable(line)) continue;
            Cursor cur(line);
            if (!cur.word(key)) continue;
            if (key == "v") { P3 p; p.x = cur.real(); p.y = cur.real(); p.z = cur.real(); vertices.push_back(p); }
            else if (key == "vn") { P3 p; p.x = cur.real(); p.y = cur.real(); p.z = cur.real(); normals.push_back(p); }
            else if (key == "o") current_mtl.clear();                            // src/Scene.cpp:268-273: the id survives
            else if (key == "usemtl") { cur.word(current_mtl); mat_id = lookup(current_mtl); }   // unknown name -> material 0 (:284-287)
            else if (key == "f") {
                if (mat_id < 0) mat_id = lookup(current_mtl);                   // faces before any usemtl (:291-297)
                TriangleCPU tri;
                std::memset(&tri.props, 0, sizeof(tri.props));                  // the reference leaves pad / missing normals indeterminate
                tri.props.matID = mat_id;
                Float4* pos[3] = {&tri.props.v1, &tri.props.v2, &tri.props.v3};
                Float4* nrm[3] = {&tri.props.vn1, &tri.props.vn2, &tri.props.vn3};
                std::string corner;
                for (int c = 0; c < 3; c++) {                                   // exactly three corners are read (:303-346)
                    if (!cur.word(corner)) throw std::runtime_error("Bad face statement in object file (fewer than 3 corners).");
                    // corner = v[/vt[/vn]] ; 1-based global indices; the vt field is skipped
                    size_t s1 = corner.find('/');
                    size_t s2 = (s1 == std::string::npos) ? std::string::npos : corner.find('/', s1 + 1);
                    long vi = std::strtol(corner.substr(0, s1).c_str(), nullptr, 10);
                    if (vi < 1 || (size_t)vi > vertices.size()) throw std::runtime_error("Bad vertex index in object file.");
                    const P3& v = vertices[vi - 1];
                    *pos[c] = {{v.x, v.y, v.z, 1.0f}};
                    if (s2 != std::string::npos && s2 + 1 < corner.size()) {
                        long ni = std::strtol(corner.c_str() + s2 + 1, nullptr, 10);
                        if (ni < 1 || (size_t)ni > normals.size()) throw std::runtime_error("Bad normal index in object file.");
                        const P3& n = normals[ni - 1];
                        *nrm[c] = {{n.x, n.y, n.z, 0.0f}};
                    }
                }
                tri.computeCentroid();
                for (int k = 0; k < 3; k++) {
                    root.p_min.s[k] = std::min(root.p_min.s[k], tri.aabb.p_min.s[k]);
                    root.p_max.s[k] = std::max(root.p_max.s[k], tri.aabb.p_max.s[k]);
                }
                cpu_tri_list.push_back(tri);
            }
            // mtllib, vt, s, g, ...: ignored
        }

        vert_data.reserve(cpu_tri_list.size());
        for (const TriangleCPU& t : cpu_tri_list) vert_data.push_back(t.props);
        if (bvh.bins > 0) bvh.createBVH(root, cpu_tri_list);      // default 20 bins, as loadModel does (:364-370)

        scene_file = filename; mat_file = mat_fp; mat_filename = mat_fn;
        num_triangles = (int)vert_data.size();
        scene_size_kb = (float)vert_data.size() * sizeof(TriangleGPU) / 1024;
        scene_size_kb += (float)mat_data.size() * sizeof(Material) / 1024;
        scene_size_mb = scene_size_kb / 1024;
    }

    void Scene::setGeometry(const TriangleGPU* tris, int n_tris, const Material* mats, int n_mats, int bvh_bins)
    {
        clearValues();
        if (n_mats <= 0 || !mats) throw std::runtime_error("Bad Material file.");
        mat_data.assign(mats, mats + n_mats);
        const float inf = std::numeric_limits<float>::max();
        root.p_min = {{inf, inf, inf, 1.0f}};
        root.p_max = {{-inf, -inf, -inf, 1.0f}};
        cpu_tri_list.resize(n_tris);
        for (int i = 0; i < n_tris; i++) {
            TriangleCPU& t = cpu_tri_list[i];
            t.props = tris[i];
            t.computeCentroid();
            for (int k = 0; k < 3; k++) {
                root.p_min.s[k] = std::min(root.p_min.s[k], t.aabb.p_min.s[k]);
                root.p_max.s[k] = std::max(root.p_max.s[k], t.aabb.p_max.s[k]);
            }
        }
        vert_data.assign(tris, tris + n_tris);
        if (bvh_bins > 0) bvh.createBVH(root, cpu_tri_list, bvh_bins);
        num_triangles = n_tris;
        scene_size_kb = (float)vert_data.size() * sizeof(TriangleGPU) / 1024 + (float)mat_data.size() * sizeof(Material) / 1024;
        scene_size_mb = scene_size_kb / 1024;
    }

    void Scene::loadBVH(int bvh_bins) { bvh.createBVH(root, cpu_tri_list, bvh_bins); }
}
