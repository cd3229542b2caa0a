// pcd_io.hpp -- ROS/PCL-free PCD v0.7 reader (ASCII, binary, binary_compressed) for the CLI harness.
// Same semantics as pcl::io::loadPCDFile as used by the reference client (src/calc_grasppoints_action_client.cpp:141):
// exactly POINTS records (extra ASCII lines ignored), tokens parsed straight to float (strtof),
// binary_compressed = uint32 compressed size, uint32 uncompressed size, LZF stream, fields stored struct-of-arrays.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <string>
#include <vector>

namespace hafpcd {

inline bool lzf_decompress(const unsigned char* ip, size_t in_len, unsigned char* out, size_t out_len) {
    size_t i = 0, o = 0;
    while (i < in_len) {
        unsigned ctrl = ip[i++];
        if (ctrl < 32) {
            size_t n = ctrl + 1;
            if (o + n > out_len || i + n > in_len) return false;
            memcpy(out + o, ip + i, n);
            i += n; o += n;
        } else {
            size_t len = ctrl >> 5;
            if (len == 7) { if (i >= in_len) return false; len += ip[i++]; }
            if (i >= in_len) return false;
            size_t dist = ((ctrl & 0x1f) << 8) + ip[i++] + 1;
            len += 2;
            if (dist > o || o + len > out_len) return false;
            for (size_t k = 0; k < len; k++, o++) out[o] = out[o - dist];
        }
    }
    return o == out_len;
}

// xyz: packed x,y,z floats.  Returns false with *err set on failure.
inline bool read_pcd(const std::string& path, std::vector<float>& xyz, std::string* err) {
    FILE* fp = fopen(path.c_str(), "rb");
    if (!fp) { if (err) *err = "cannot open " + path; return false; }
    std::vector<unsigned char> raw;
    unsigned char buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, fp)) > 0) raw.insert(raw.end(), buf, buf + n);
    fclose(fp);
    size_t pos = 0;
    std::vector<std::string> fields, types;
    std::vector<int> sizes, counts;
    long npts = -1, width = 0, height = 1;
    std::string data_kind;
    while (pos < raw.size()) {
        size_t nl = pos;
        while (nl < raw.size() && raw[nl] != '\n') nl++;
        std::string line((const char*)&raw[pos], nl - pos);
        pos = nl + 1;
        if (line.empty() || line[0] == '#') continue;
        std::istringstream ss(line);
        std::string key;
        ss >> key;
        std::string tok;
        if (key == "FIELDS") while (ss >> tok) fields.push_back(tok);
        else if (key == "SIZE") while (ss >> tok) sizes.push_back(atoi(tok.c_str()));
        else if (key == "TYPE") while (ss >> tok) types.push_back(tok);
        else if (key == "COUNT") while (ss >> tok) counts.push_back(atoi(tok.c_str()));
        else if (key == "WIDTH") ss >> width;
        else if (key == "HEIGHT") ss >> height;
        else if (key == "POINTS") ss >> npts;
        else if (key == "DATA") { ss >> data_kind; break; }
    }
    if (npts < 0) npts = width * height;
    if (counts.empty()) counts.assign(fields.size(), 1);
    int ix = -1, iy = -1, iz = -1;
    for (size_t k = 0; k < fields.size(); k++) { if (fields[k] == "x") ix = (int)k; if (fields[k] == "y") iy = (int)k; if (fields[k] == "z") iz = (int)k; }
    if (ix < 0 || iy < 0 || iz < 0 || sizes.size() != fields.size() || types.size() != fields.size()) { if (err) *err = "PCD header lacks x/y/z fields"; return false; }
    xyz.assign((size_t)npts * 3, 0.0f);
    const int idx[3] = {ix, iy, iz};
    if (data_kind == "ascii") {
        std::vector<int> tokoff(fields.size() + 1, 0);
        for (size_t k = 0; k < fields.size(); k++) tokoff[k + 1] = tokoff[k] + counts[k];
        long rec = 0;
        while (pos < raw.size() && rec < npts) {
            size_t nl = pos;
            while (nl < raw.size() && raw[nl] != '\n') nl++;
            std::string line((const char*)&raw[pos], nl - pos);
            pos = nl + 1;
            std::vector<std::string> toks;
            std::istringstream ss(line);
            std::string tok;
            while (ss >> tok) toks.push_back(tok);
            if (toks.empty()) continue;
            for (int a = 0; a < 3; a++) {
                if (tokoff[idx[a]] >= (int)toks.size()) { if (err) *err = "short ASCII record in " + path; return false; }
                xyz[rec * 3 + a] = strtof(toks[tokoff[idx[a]]].c_str(), NULL);
            }
            rec++;
        }
        if (rec != npts) { if (err) *err = "fewer records than POINTS in " + path; return false; }
        return true;
    }
    for (int a = 0; a < 3; a++) if (types[idx[a]] != "F" || sizes[idx[a]] != 4) { if (err) *err = "x/y/z must be float32"; return false; }
    if (data_kind == "binary") {
        size_t rec_bytes = 0;
        std::vector<size_t> foff(fields.size());
        for (size_t k = 0; k < fields.size(); k++) { foff[k] = rec_bytes; rec_bytes += (size_t)sizes[k] * counts[k]; }
        if (pos + rec_bytes * npts > raw.size()) { if (err) *err = "binary PCD truncated"; return false; }
        for (long r = 0; r < npts; r++)
            for (int a = 0; a < 3; a++) memcpy(&xyz[r * 3 + a], &raw[pos + r * rec_bytes + foff[idx[a]]], 4);
        return true;
    }
    if (data_kind == "binary_compressed") {
        if (pos + 8 > raw.size()) { if (err) *err = "compressed PCD truncated"; return false; }
        uint32_t comp, uncomp;
        memcpy(&comp, &raw[pos], 4);
        memcpy(&uncomp, &raw[pos + 4], 4);
        if (pos + 8 + comp > raw.size()) { if (err) *err = "compressed PCD truncated"; return false; }
        std::vector<unsigned char> blob(uncomp);
        if (!lzf_decompress(&raw[pos + 8], comp, blob.data(), uncomp)) { if (err) *err = "corrupt LZF stream in " + path; return false; }
        std::vector<size_t> foff(fields.size());
        size_t off = 0;
        for (size_t k = 0; k < fields.size(); k++) { foff[k] = off; off += (size_t)npts * sizes[k] * counts[k]; }
        for (int a = 0; a < 3; a++)
            for (long r = 0; r < npts; r++) memcpy(&xyz[r * 3 + a], &blob[foff[idx[a]] + (size_t)r * 4 * counts[idx[a]]], 4);
        return true;
    }
    if (err) *err = "unsupported PCD DATA kind " + data_kind;
    return false;
}

}  // namespace hafpcd
