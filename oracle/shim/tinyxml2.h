// TEST INFRASTRUCTURE ONLY (oracle tier A). Stand-in for the un-vendored tinyxml2 dependency of the
// reference (scene.h:8, scene.cpp:5-53): XMLDocument::LoadFile/ErrorStr/RootElement and
// XMLElement::Attribute/FirstChildElement/NextSiblingElement.  Like tinyxml2 it accepts several
// top-level elements (cornell-box.xml:2-7) and keeps newlines inside attribute values
// (staircase.xml:10-12).  No entities, CDATA or DTDs: the cg22 scene files have none.
#pragma once
#include <cctype>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace tinyxml2
{
enum XMLError
{
    XML_SUCCESS = 0,
    XML_ERROR_FILE_NOT_FOUND,
    XML_ERROR_PARSING
};

class XMLElement
{
public:
    const char *Name() const { return name.c_str(); }
    const char *Attribute(const char *key) const
    {
        for (auto &kv : attrs)
            if (kv.first == key)
                return kv.second.c_str();
        return nullptr;
    }
    XMLElement *FirstChildElement(const char *n = nullptr)
    {
        for (auto &c : children)
            if (!n || c->name == n)
                return c.get();
        return nullptr;
    }
    XMLElement *NextSiblingElement(const char *n = nullptr)
    {
        if (!siblings)
            return nullptr;
        for (size_t i = index + 1; i < siblings->size(); i++)
            if (!n || (*siblings)[i]->name == n)
                return (*siblings)[i].get();
        return nullptr;
    }

    std::string name;
    std::vector<std::pair<std::string, std::string>> attrs;
    std::vector<std::unique_ptr<XMLElement>> children;
    std::vector<std::unique_ptr<XMLElement>> *siblings = nullptr;
    size_t index = 0;
};

class XMLDocument
{
public:
    XMLError LoadFile(const char *path)
    {
        FILE *f = fopen(path, "rb");
        if (!f)
        {
            err = "XML_ERROR_FILE_NOT_FOUND";
            return XML_ERROR_FILE_NOT_FOUND;
        }
        std::string s;
        char buf[4096];
        size_t n;
        while ((n = fread(buf, 1, sizeof buf, f)) > 0)
            s.append(buf, n);
        fclose(f);
        size_t p = 0;
        if (!parseNodes(s, p, roots, nullptr))
        {
            err = "XML_ERROR_PARSING";
            return XML_ERROR_PARSING;
        }
        return XML_SUCCESS;
    }
    const char *ErrorStr() const { return err.c_str(); }
    XMLElement *RootElement() { return roots.empty() ? nullptr : roots[0].get(); }

private:
    static void skipWs(const std::string &s, size_t &p)
    {
        while (p < s.size() && isspace((unsigned char)s[p]))
            p++;
    }
    // parses siblings until the closing tag of `parent` (or end of input when parent == nullptr)
    static bool parseNodes(const std::string &s, size_t &p, std::vector<std::unique_ptr<XMLElement>> &out,
                           const XMLElement *parent)
    {
        for (;;)
        {
            size_t lt = s.find('<', p);
            if (lt == std::string::npos)
                return parent == nullptr;
            p = lt;
            if (s.compare(p, 2, "<?") == 0)
            {
                size_t e = s.find("?>", p);
                if (e == std::string::npos)
                    return false;
                p = e + 2;
                continue;
            }
            if (s.compare(p, 4, "<!--") == 0)
            {
                size_t e = s.find("-->", p);
                if (e == std::string::npos)
                    return false;
                p = e + 3;
                continue;
            }
            if (s.compare(p, 2, "</") == 0)
            {
                size_t e = s.find('>', p);
                if (e == std::string::npos || !parent)
                    return false;
                p = e + 1;
                return true;
            }
            p++; // '<'
            std::unique_ptr<XMLElement> el(new XMLElement());
            while (p < s.size() && !isspace((unsigned char)s[p]) && s[p] != '>' && s[p] != '/')
                el->name.push_back(s[p++]);
            bool selfClosed = false;
            for (;;)
            {
                skipWs(s, p);
                if (p >= s.size())
                    return false;
                if (s[p] == '/')
                {
                    selfClosed = true;
                    p++;
                    continue;
                }
                if (s[p] == '>')
                {
                    p++;
                    break;
                }
                std::string key, val;
                while (p < s.size() && s[p] != '=' && !isspace((unsigned char)s[p]))
                    key.push_back(s[p++]);
                skipWs(s, p);
                if (p >= s.size() || s[p] != '=')
                    return false;
                p++;
                skipWs(s, p);
                if (p >= s.size() || (s[p] != '"' && s[p] != '\''))
                    return false;
                char q = s[p++];
                while (p < s.size() && s[p] != q)
                {
                    if (s[p] == '\r')
                    { // newline normalisation as tinyxml2 (CRLF / CR -> LF)
                        val.push_back('\n');
                        if (p + 1 < s.size() && s[p + 1] == '\n')
                            p++;
                        p++;
                        continue;
                    }
                    val.push_back(s[p++]);
                }
                if (p >= s.size())
                    return false;
                p++;
                el->attrs.emplace_back(key, val);
            }
            el->siblings = &out;
            el->index = out.size();
            XMLElement *raw = el.get();
            out.push_back(std::move(el));
            if (!selfClosed && !parseNodes(s, p, raw->children, raw))
                return false;
        }
    }

    std::vector<std::unique_ptr<XMLElement>> roots;
    std::string err;
};
} // namespace tinyxml2
