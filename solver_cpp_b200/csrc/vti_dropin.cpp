// The two geometry-file symbols the reference's DiscretizationGrid binds (/root/reference/src/grid/discretization_grid.h:20-22):
//   f_VTI_readCellsSizeOrigin   /root/reference/src/grid/VTI.f90:146-266   cells from WholeExtent, size = Spacing * cells, Origin
//   f_VTI_readDataset_int       /root/reference/src/grid/VTI.f90:473-497   a cell-data array as 32-bit integers, PLUS ONE (:494)
// for VTK ImageData files as DAMASK writes them (python/damask/_grid.py: base64 payload, optional vtkZLibDataCompressor blocks,
// UInt32 / UInt64 headers, ascii arrays).  solver_cpp_b200/vti.py is the Python host's reader of the same files.
// Errors cannot be returned through these signatures: cells / size are left at -1 and the message is in dmk_last_error().
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include <zlib.h>

#include "dmk_common.cuh"

namespace {

[[noreturn]] void bad(const std::string& what) { throw std::runtime_error("VTI: " + what); }

// the first <name ...> tag at or after `from`; returns its text without the brackets, and its end position
bool find_tag(const std::string& s, const std::string& name, size_t from, std::string* tag, size_t* end) {
  size_t p = from;
  for (;;) {
    p = s.find("<" + name, p);
    if (p == std::string::npos) return false;
    const char next = s[p + 1 + name.size()];
    if (next == ' ' || next == '>' || next == '\n' || next == '\t' || next == '\r' || next == '/') break;
    p++;
  }
  const size_t q = s.find('>', p);
  if (q == std::string::npos) return false;
  *tag = s.substr(p + 1, q - p - 1);
  *end = q + 1;
  return true;
}
bool attr(const std::string& tag, const std::string& name, std::string* value) {
  size_t p = 0;
  for (;;) {
    p = tag.find(name, p);
    if (p == std::string::npos) return false;
    const bool starts = p == 0 || tag[p - 1] == ' ' || tag[p - 1] == '\n' || tag[p - 1] == '\t';
    size_t q = p + name.size();
    while (q < tag.size() && (tag[q] == ' ')) q++;
    if (starts && q < tag.size() && tag[q] == '=') {
      q = tag.find('"', q);
      if (q == std::string::npos) return false;
      const size_t e = tag.find('"', q + 1);
      if (e == std::string::npos) return false;
      *value = tag.substr(q + 1, e - q - 1);
      return true;
    }
    p++;
  }
}
std::string need_attr(const std::string& tag, const std::string& name) {
  std::string v;
  if (!attr(tag, name, &v)) bad("attribute " + name + " missing in <" + tag.substr(0, tag.find(' ')) + ">");
  return v;
}

std::vector<unsigned char> b64decode(const std::string& text, size_t begin, size_t count) {
  static int T[256];
  static bool init = false;
  if (!init) {
    for (int i = 0; i < 256; i++) T[i] = -1;
    const char* A = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
    for (int i = 0; i < 64; i++) T[(unsigned char)A[i]] = i;
    init = true;
  }
  std::vector<unsigned char> out;
  out.reserve(count / 4 * 3);
  unsigned acc = 0; int bits = 0;
  for (size_t i = begin; i < begin + count && i < text.size(); i++) {
    const unsigned char ch = (unsigned char)text[i];
    if (ch == '=') break;
    if (T[ch] < 0) bad("invalid base64 character");
    acc = (acc << 6) | (unsigned)T[ch]; bits += 6;
    if (bits >= 8) { bits -= 8; out.push_back((unsigned char)((acc >> bits) & 0xffu)); }
  }
  return out;
}
size_t b64len(size_t n_bytes) { return 4 * ((n_bytes + 2) / 3); }

size_t type_size(const std::string& t) {
  if (t == "Int8" || t == "UInt8") return 1;
  if (t == "Int16" || t == "UInt16") return 2;
  if (t == "Int32" || t == "UInt32" || t == "Float32") return 4;
  if (t == "Int64" || t == "UInt64" || t == "Float64") return 8;
  bad("unknown data type " + t);
}
uint64_t header_value(const unsigned char* p, size_t hsize) {
  uint64_t v = 0;
  memcpy(&v, p, hsize);                                        // little endian (checked by the caller)
  return v;
}
long long element(const unsigned char* p, const std::string& t) {
  if (t == "Int8") { int8_t v; memcpy(&v, p, 1); return v; }
  if (t == "UInt8") { uint8_t v; memcpy(&v, p, 1); return v; }
  if (t == "Int16") { int16_t v; memcpy(&v, p, 2); return v; }
  if (t == "UInt16") { uint16_t v; memcpy(&v, p, 2); return v; }
  if (t == "Int32") { int32_t v; memcpy(&v, p, 4); return v; }
  if (t == "UInt32") { uint32_t v; memcpy(&v, p, 4); return v; }
  if (t == "Int64") { int64_t v; memcpy(&v, p, 8); return v; }
  if (t == "UInt64") { uint64_t v; memcpy(&v, p, 8); return (long long)v; }
  if (t == "Float32") { float v; memcpy(&v, p, 4); return (long long)v; }
  if (t == "Float64") { double v; memcpy(&v, p, 8); return (long long)v; }
  bad("unknown data type " + t);
}

struct FileInfo { std::string header_type; bool compressed; std::string image_tag; };
FileInfo file_info(const std::string& txt) {
  std::string vtk; size_t end = 0;
  if (!find_tag(txt, "VTKFile", 0, &vtk, &end)) bad("no <VTKFile>");
  if (need_attr(vtk, "type") != "ImageData") bad("not an ImageData file");
  if (need_attr(vtk, "byte_order") != "LittleEndian") bad("only little-endian files are supported");
  FileInfo f;
  if (!attr(vtk, "header_type", &f.header_type)) f.header_type = "UInt32";
  std::string comp;
  f.compressed = attr(vtk, "compressor", &comp);
  if (f.compressed && comp != "vtkZLibDataCompressor") bad("unsupported compressor " + comp);
  if (!find_tag(txt, "ImageData", end, &f.image_tag, &end)) bad("no <ImageData>");
  return f;
}
std::vector<double> numbers(const std::string& s) {
  std::vector<double> v;
  const char* p = s.c_str();
  char* e = nullptr;
  for (;;) { const double x = strtod(p, &e); if (e == p) break; v.push_back(x); p = e; }
  return v;
}
void cells_size_origin(const FileInfo& f, int cells[3], double size[3], double origin[3]) {
  const std::vector<double> ext = numbers(need_attr(f.image_tag, "WholeExtent")), sp = numbers(need_attr(f.image_tag, "Spacing")),
                            org = numbers(need_attr(f.image_tag, "Origin"));
  if (ext.size() != 6 || sp.size() != 3 || org.size() != 3) bad("malformed WholeExtent / Spacing / Origin");
  for (int d = 0; d < 3; d++) {
    cells[d] = (int)(ext[2 * d + 1] - ext[2 * d]);
    size[d] = sp[d] * cells[d];
    origin[d] = org[d];
  }
  if (cells[0] < 1 || cells[1] < 1 || cells[2] < 1) bad("invalid cell counts");
}

std::vector<long long> read_dataset(const std::string& txt, const std::string& label) {
  const FileInfo f = file_info(txt);
  int cells[3]; double size[3], origin[3];
  cells_size_origin(f, cells, size, origin);
  const size_t n = (size_t)cells[0] * cells[1] * cells[2];
  const size_t cd = txt.find("<CellData");
  const size_t cd_end = txt.find("</CellData>");
  if (cd == std::string::npos || cd_end == std::string::npos) bad("no CellData");
  size_t pos = cd;
  for (;;) {
    std::string tag; size_t body = 0;
    if (!find_tag(txt, "DataArray", pos, &tag, &body) || body > cd_end) bad("cell dataset " + label + " not found");
    const size_t close = txt.find("</DataArray>", body);
    if (close == std::string::npos) bad("unterminated DataArray");
    pos = close + 12;
    std::string name;
    if (!attr(tag, "Name", &name) || name != label) continue;
    const std::string type = need_attr(tag, "type"), format = need_attr(tag, "format");
    std::vector<long long> out;
    if (format == "ascii") {
      for (double v : numbers(txt.substr(body, close - body))) out.push_back((long long)v);
    } else if (format == "binary") {
      std::string payload;
      for (size_t i = body; i < close; i++) { const char ch = txt[i]; if (ch != ' ' && ch != '\n' && ch != '\r' && ch != '\t') payload += ch; }
      const size_t hs = type_size(f.header_type), es = type_size(type);
      std::vector<unsigned char> raw;
      if (!f.compressed) {
        const std::vector<unsigned char> all = b64decode(payload, 0, payload.size());
        if (all.size() < hs) bad("truncated payload");
        const size_t nbytes = (size_t)header_value(all.data(), hs);
        if (all.size() < hs + nbytes) bad("truncated payload");
        raw.assign(all.begin() + (long)hs, all.begin() + (long)(hs + nbytes));
      } else {
        const std::vector<unsigned char> first = b64decode(payload, 0, b64len(3 * hs));
        if (first.size() < 3 * hs) bad("truncated block header");
        const size_t n_blocks = (size_t)header_value(first.data(), hs), block = (size_t)header_value(first.data() + hs, hs),
                     last = (size_t)header_value(first.data() + 2 * hs, hs);
        const size_t hlen = b64len((3 + n_blocks) * hs);
        const std::vector<unsigned char> head = b64decode(payload, 0, hlen);
        if (head.size() < (3 + n_blocks) * hs) bad("truncated block header");
        const std::vector<unsigned char> data = b64decode(payload, hlen, payload.size() - hlen);
        size_t p = 0;
        for (size_t b = 0; b < n_blocks; b++) {
          const size_t csize = (size_t)header_value(head.data() + (3 + b) * hs, hs);
          uLongf usize = (uLongf)((b + 1 == n_blocks && last) ? last : block);
          std::vector<unsigned char> buf(usize);
          if (p + csize > data.size() || uncompress(buf.data(), &usize, data.data() + p, (uLong)csize) != Z_OK) bad("zlib block " + std::to_string(b) + " does not inflate");
          raw.insert(raw.end(), buf.begin(), buf.begin() + (long)usize);
          p += csize;
        }
      }
      if (raw.size() % es) bad("payload is not a whole number of elements");
      for (size_t i = 0; i < raw.size() / es; i++) out.push_back(element(raw.data() + i * es, type));
    } else bad("unsupported DataArray format " + format);
    if (out.size() != n) bad("dataset " + label + " has " + std::to_string(out.size()) + " entries, the grid has " + std::to_string(n) + " cells");
    return out;
  }
}

}  // namespace

extern "C" void f_VTI_readCellsSizeOrigin(const char* fileContent, int string_size, int* cells, double* size, double* origin) {
  for (int d = 0; d < 3; d++) { cells[d] = -1; size[d] = -1.0; origin[d] = 0.0; }
  try {
    const std::string txt(fileContent, (size_t)string_size);
    cells_size_origin(file_info(txt), cells, size, origin);
  } catch (const std::exception& e) {
    dmk_set_error("%s", e.what());
    for (int d = 0; d < 3; d++) { cells[d] = -1; size[d] = -1.0; }
  }
}

extern "C" void f_VTI_readDataset_int(const char* fileContent, int string_size, const char* label, int label_length, int* dataset_ptr,
                                      int* dataset_size) {
  try {
    const std::vector<long long> v = read_dataset(std::string(fileContent, (size_t)string_size), std::string(label, (size_t)label_length));
    if (dataset_size && *dataset_size > 0 && (size_t)*dataset_size < v.size()) throw std::runtime_error("VTI: the caller's buffer is too small");
    for (size_t i = 0; i < v.size(); i++) dataset_ptr[i] = (int)v[i] + 1;               // VTI.f90:494: material IDs become 1-based
    if (dataset_size) *dataset_size = (int)v.size();
  } catch (const std::exception& e) {
    dmk_set_error("%s", e.what());
    if (dataset_size) *dataset_size = -1;
  }
}
