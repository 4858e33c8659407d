// Scene loaders: XML (camera + lights), MTL, OBJ — host side, cold path.
// Behaviour follows the reference's hand-rolled loaders including their quirks:
//   scene.cpp:3-55    readxml: first top-level element = camera, following <light> siblings, radiance split
//                     at the first two commas, multi-line attribute values
//   scene.cpp:57-113  readmtl: only newmtl/Kd/Ks/Tr/Ns/Ni/map_Kd are read (`Kt` is ignored)
//   scene.cpp:115-213 readobj: triangles only (first three corners), `isvnvt` slot order, derived fields
//                     normal / center / cumulative light area computed in the same float order
// Errors throw trt::LoadError instead of exit().
#include "tinyrt.h"

#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>

namespace trt
{
namespace
{
struct XmlElem
{
    int depth;
    std::string name;
    std::vector<std::pair<std::string, std::string>> attrs;
    const std::string *get(const char *k) const
    {
        for (auto &a : attrs)
            if (a.first == k)
                return &a.second;
        return nullptr;
    }
};

// Flat scan of start tags in document order with their nesting depth. Several top-level elements are fine.
std::vector<XmlElem> scanXml(const std::string &s)
{
    std::vector<XmlElem> out;
    int depth = 0;
    size_t i = 0;
    auto ws = [&] { while (i < s.size() && std::isspace((unsigned char)s[i])) ++i; };
    while ((i = s.find('<', i)) != std::string::npos)
    {
        if (s.compare(i, 2, "<?") == 0)
        {
            i = s.find("?>", i);
            if (i == std::string::npos)
                throw LoadError("xml: unterminated declaration");
            continue;
        }
        if (s.compare(i, 4, "<!--") == 0)
        {
            i = s.find("-->", i);
            if (i == std::string::npos)
                throw LoadError("xml: unterminated comment");
            continue;
        }
        if (s.compare(i, 2, "</") == 0)
        {
            --depth;
            ++i;
            continue;
        }
        ++i;
        XmlElem e;
        e.depth = depth;
        while (i < s.size() && !std::isspace((unsigned char)s[i]) && s[i] != '>' && s[i] != '/')
            e.name += s[i++];
        bool closed = false;
        for (;;)
        {
            ws();
            if (i >= s.size())
                throw LoadError("xml: unterminated tag <" + e.name);
            if (s[i] == '>')
            {
                ++i;
                break;
            }
            if (s[i] == '/')
            {
                closed = true;
                ++i;
                continue;
            }
            std::string k, val;
            while (i < s.size() && s[i] != '=' && !std::isspace((unsigned char)s[i]))
                k += s[i++];
            ws();
            if (i >= s.size() || s[i] != '=')
                throw LoadError("xml: attribute without value in <" + e.name);
            ++i;
            ws();
            if (i >= s.size() || (s[i] != '"' && s[i] != '\''))
                throw LoadError("xml: unquoted attribute in <" + e.name);
            const char q = s[i++];
            while (i < s.size() && s[i] != q)
            {
                if (s[i] == '\r')
                { // CRLF / CR -> LF, as tinyxml2 normalises attribute text
                    val += '\n';
                    i += (i + 1 < s.size() && s[i + 1] == '\n') ? 2 : 1;
                }
                else
                    val += s[i++];
            }
            ++i;
            e.attrs.emplace_back(k, val);
        }
        out.push_back(e);
        if (!closed)
            ++depth;
    }
    return out;
}

float toFloat(const std::string &s, const char *what)
{
    const char *b = s.c_str();
    char *e = nullptr;
    float f = std::strtof(b, &e); // std::stof semantics: leading whitespace skipped, trailing text ignored
    if (e == b)
        throw LoadError(std::string("not a number: ") + what + " = '" + s + "'");
    return f;
}

const std::string &need(const XmlElem &e, const char *k)
{
    const std::string *p = e.get(k);
    if (!p)
        throw LoadError("xml: <" + e.name + "> lacks attribute " + k);
    return *p;
}

vec3 xyz(const XmlElem &e) { return vec3(toFloat(need(e, "x"), "x"), toFloat(need(e, "y"), "y"), toFloat(need(e, "z"), "z")); }
} // namespace

void Scene::readxml(std::string xml_path)
{
    std::ifstream fin(xml_path, std::ios::binary);
    if (!fin.is_open())
        throw LoadError("Read xml failed: " + xml_path);
    std::stringstream ss;
    ss << fin.rdbuf();
    std::vector<XmlElem> el = scanXml(ss.str());
    if (el.empty())
        throw LoadError("Read xml failed: no element in " + xml_path);
    const XmlElem &cam = el[0]; // RootElement(): the first top-level element, whatever its name
    img_width = std::stoi(need(cam, "width"));
    img_height = std::stoi(need(cam, "height"));
    camera.aspect_ratio = (double)img_width / (double)img_height;
    camera.fovy = toFloat(need(cam, "fovy"), "fovy"); // stof -> float -> double (scene.cpp:16)
    bool haveEye = false, haveLook = false, haveUp = false;
    size_t k = 1;
    for (; k < el.size() && el[k].depth > 0; ++k)
    {
        if (el[k].depth != 1)
            continue;
        if (el[k].name == "eye" && !haveEye)
            camera.eye = xyz(el[k]), haveEye = true;
        else if (el[k].name == "lookat" && !haveLook)
            camera.lookat = xyz(el[k]), haveLook = true;
        else if (el[k].name == "up" && !haveUp)
            camera.up = xyz(el[k]), haveUp = true;
    }
    if (!haveEye || !haveLook || !haveUp)
        throw LoadError("xml: camera needs eye, lookat and up");
    camera.setCamera();
    for (; k < el.size(); ++k)
    {
        if (el[k].depth != 0 || el[k].name != "light")
            continue;
        const std::string name = need(el[k], "mtlname");
        const std::string &rs = need(el[k], "radiance");
        // scene.cpp:31-49: x up to the first comma, y up to the second, z = the rest
        vec3 radiance;
        const size_t c1 = rs.find(',');
        const size_t c2 = (c1 == std::string::npos) ? std::string::npos : rs.find(',', c1 + 1);
        size_t rest = 0;
        if (c1 != std::string::npos)
            radiance.x = toFloat(rs.substr(0, c1), "radiance.x"), rest = c1 + 1;
        if (c2 != std::string::npos)
            radiance.y = toFloat(rs.substr(c1 + 1, c2 - c1 - 1), "radiance.y"), rest = c2 + 1;
        if (!rs.empty() && rs.size() - 1 != c1 && rs.size() - 1 != c2) // last char not consumed as a separator
            radiance.z = toFloat(rs.substr(rest), "radiance.z");
        lights.push_back(Light(name, radiance));
        materials[name].is_emissive = true;
        materials[name].radiance = radiance;
    }
}

void Scene::readmtl(std::string mtl_path, std::string basedir)
{
    std::ifstream fin(mtl_path);
    if (!fin.is_open())
        throw LoadError("Read " + mtl_path + " failed.");
    std::string line, cur;
    while (std::getline(fin, line))
    {
        std::istringstream in(line);
        std::string key;
        in >> key;
        float x = 0, y = 0, z = 0;
        if (key == "newmtl")
            in >> cur;
        else if (key == "Kd" || key == "Ks" || key == "Tr")
        {
            in >> x >> y >> z;
            Material &m = materials[cur];
            (key == "Kd" ? m.Kd : key == "Ks" ? m.Ks : m.Tr) = vec3(x, y, z);
        }
        else if (key == "Ns" || key == "Ni")
        {
            in >> x;
            (key == "Ns" ? materials[cur].Ns : materials[cur].Ni) = x;
        }
        else if (key == "map_Kd")
        {
            std::string rel;
            in >> rel;
            materials[cur].map_Kd = basedir + "/" + rel;
            materials[cur].readinMap();
        }
        // everything else (Kt, Ka, illum, ...) is ignored, as in the reference
    }
    std::printf("num of materials: %d\n", (int)materials.size());
}

void Scene::readobj(std::string obj_path)
{
    std::ifstream fin(obj_path);
    if (!fin.is_open())
        throw LoadError("Read " + obj_path + " failed.");
    std::vector<vec3> pos, nrm;
    std::vector<vec2> uv;
    bool slot2_is_normal = true; // the reference's `isvnvt`: cleared by a vt line seen before any vn line
    std::string line, mtl_name;
    int face = 0;
    auto pick = [&](const std::string &tok, size_t size, const char *what) -> int {
        int idx;
        try
        {
            idx = std::stoi(tok) - 1;
        }
        catch (...)
        {
            throw LoadError("obj: bad index '" + tok + "' in face " + std::to_string(face));
        }
        if (idx < 0 || (size_t)idx >= size)
            throw LoadError(std::string("obj: ") + what + " index out of range in face " + std::to_string(face));
        return idx;
    };
    while (std::getline(fin, line))
    {
        std::istringstream in(line);
        std::string key;
        in >> key;
        float x = 0, y = 0, z = 0;
        if (key == "v")
        {
            in >> x >> y >> z;
            pos.push_back(vec3(x, y, z));
        }
        else if (key == "vn")
        {
            in >> x >> y >> z;
            nrm.push_back(vec3(x, y, z));
        }
        else if (key == "vt")
        {
            if (nrm.empty())
                slot2_is_normal = false;
            in >> x >> y;
            uv.push_back(vec2(x, y));
        }
        else if (key == "usemtl")
            in >> mtl_name;
        else if (key == "f")
        {
            Triangle t;
            for (int c = 0; c < 3; ++c) // only the first three corners are read (scene.cpp:162)
            {
                std::string tok;
                in >> tok;
                if (tok.empty())
                    continue;
                // fields between '/': first (when a '/' follows) -> position; last -> third slot; the ones
                // in between -> second slot; a token without '/' only fills the third slot (scene.cpp:167-193)
                std::vector<std::string> f;
                size_t b = 0, p;
                while ((p = tok.find('/', b)) != std::string::npos)
                {
                    f.push_back(tok.substr(b, p - b));
                    b = p + 1;
                }
                f.push_back(tok.substr(b));
                for (size_t q = 0; q < f.size(); ++q)
                {
                    const bool last = (q + 1 == f.size());
                    if (q == 0 && !last)
                        t.v[c] = pos[pick(f[q], pos.size(), "v")];
                    else if ((!last) == slot2_is_normal) // middle+isvnvt or last+!isvnvt -> normal
                        t.vn[c] = nrm[pick(f[q], nrm.size(), "vn")];
                    else
                        t.vt[c] = uv[pick(f[q], uv.size(), "vt")];
                }
            }
            t.normal = normalize(cross(t.v[1] - t.v[0], t.v[2] - t.v[0]));
            t.center = (t.v[0] + t.v[1] + t.v[2]) / vec3(3.0f);
            t.mtl_name = mtl_name;
            t.face = face++;
            Material &m = materials[mtl_name];
            if (m.is_emissive)
            {
                t.is_emissive = true;
                m.area += t.calAera();
                t.area = m.area;
                m.triangles.push_back(t);
            }
            triangles.push_back(t);
        }
    }
    std::printf("num of vertices: %d\nnum of vn: %d\nnum of vt: %d\nnum of triangles: %d\n", (int)pos.size(),
                (int)nrm.size(), (int)uv.size(), (int)triangles.size());
}

// material.cpp:3-11 — the reference calls cv::imread.  Baseline JPEG (all cg22 textures) is decoded here, bit for bit
// as OpenCV's decoder does (jpeg_decoder.cpp); anything else is read from a pre-decoded side-car "<file>.bgr"
// ("BGR8", int32 rows, int32 cols, bytes) written with cv2.imread by tinyraytracing_b200.scenes.write_bgr_sidecar.
void Material::readinMap()
{
    std::printf("Reading map_Kd file %s\n", map_Kd.c_str());
    img = Image();
    std::string why;
    if (!decodeJpegFile(map_Kd, img, why))
    {
        img = Image();
        if (FILE *f = std::fopen((map_Kd + ".bgr").c_str(), "rb"))
        {
            char magic[4];
            int32_t rc[2];
            if (std::fread(magic, 1, 4, f) == 4 && std::string(magic, 4) == "BGR8" && std::fread(rc, 4, 2, f) == 2 &&
                rc[0] > 0 && rc[1] > 0)
            {
                auto buf = std::make_shared<std::vector<unsigned char>>((size_t)rc[0] * rc[1] * 3);
                if (std::fread(buf->data(), 1, buf->size(), f) == buf->size())
                {
                    img.data = buf;
                    img.rows = rc[0];
                    img.cols = rc[1];
                }
            }
            std::fclose(f);
        }
    }
    if (img.empty())
        std::printf("Cannot read file: %s (%s; no side-car %s.bgr either)\n", map_Kd.c_str(), why.c_str(), map_Kd.c_str());
    map_height = img.rows;
    map_width = img.cols;
}

// triangle.cpp:3-10
double Triangle::calAera()
{
    double a = length(v[1] - v[0]), b = length(v[2] - v[0]), c = length(v[2] - v[1]);
    double cos_c = (a * a + b * b - c * c) / (2 * a * b);
    double sin_c = std::sqrt(1 - std::pow(cos_c, 2));
    return a * b * sin_c / 2;
}

// camera.cpp:3-17
void Camera::setCamera()
{
    double theta = fovy * 0.01745329251994329576923690768489; // glm::radians<double>
    double h = std::tan(theta / 2);
    float viewport_height = (float)(2.0 * h);
    float viewport_width = (float)(aspect_ratio * viewport_height);
    vec3 w = normalize(eye - lookat);
    vec3 u = normalize(cross(up, w));
    vec3 v = cross(w, u);
    horizontal = viewport_width * u;
    vertical = viewport_height * v;
    lower_left_corner = eye - horizontal / 2.0f - vertical / 2.0f - w;
}

// camera.cpp:19-28
Ray Camera::getRay(float s, float t)
{
    Ray ray;
    ray.startpoint = eye;
    ray.direction = normalize(lower_left_corner + s * horizontal + t * vertical - eye);
    return ray;
}

void Camera::Print()
{
    std::printf("Camera:\nfovy: %f eye: (%f, %f, %f) lookat: (%f, %f, %f) up: (%f, %f, %f) \n", fovy, eye.x, eye.y,
                eye.z, lookat.x, lookat.y, lookat.z, up.x, up.y, up.z);
}
} // namespace trt
