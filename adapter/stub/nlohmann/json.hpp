// compile-check stub, see stub/README.md (only the members CUDASolver.cpp uses)
#pragma once
#include <map>
#include <string>
namespace nlohmann
{
    class json
    {
    public:
        using map_t = std::map<std::string, json>;
        bool contains(const std::string &k) const { return obj.count(k) != 0; }
        json &operator[](const std::string &k) { return obj[k]; }
        const json &operator[](const std::string &k) const { return obj.at(k); }
        std::string dump() const { return "{}"; }
        static json parse(const char *) { return json(); }
        struct iterator
        {
            map_t::const_iterator it;
            const std::string &key() const { return it->first; }
            const json &value() const { return it->second; }
            iterator &operator++() { ++it; return *this; }
            bool operator!=(const iterator &o) const { return it != o.it; }
        };
        iterator begin() const { return {obj.begin()}; }
        iterator end() const { return {obj.end()}; }
        map_t obj;
    };
} // namespace nlohmann
