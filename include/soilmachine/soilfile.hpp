// include/soilmachine/soilfile.hpp -- reader for the `.soil` preset format.
//
// Presets written for the reference (its loader is loadsoil(), source/io.h:7-230) must produce the same soil /
// layer / world tables here, so the reader keeps that loader's observable behaviour:
//   * everything from `#` to the end of a line is a comment; a line that is empty after that is skipped, any other
//     line needs a space between keyword and value;
//   * three block kinds, `SOIL <name> {`, `LAYER <name> {`, `WORLD {`, closed by a line holding only `}`;
//   * soil ids are handed out in order of first mention (a cross reference such as `TRANSPORTS Sand` may come
//     before `SOIL Sand {`); "Air" is id 0;
//   * soil parameters carry over from one SOIL block to the next unless a block sets them (the reference edits
//     one running record); LAYER blocks are kept in file order = deposition order;
//   * unknown keywords inside a block are ignored (e.g. a `SEED` line inside WORLD);
//   * colours are six hex digits 0-9A-F -> rgb/255, alpha 1.
// The implementation is table-driven: one descriptor per keyword names the block it belongs to, how its value
// is decoded and where the result is stored.  Errors throw SoilFileError (the reference prints and exits).
#pragma once
#include <cstddef>
#include <cstring>
#include <fstream>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace soilmachine {

struct SoilFileError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

struct SoilEntry {            // numeric + name fields of SurfParam, surface.h:11-39
  std::string name;
  float density = 0.0f, porosity = 0.0f;
  float color[4] = {0.5f, 0.5f, 0.5f, 1.0f};
  float phong[4] = {0.5f, 0.8f, 0.2f, 32.0f};
  int transports = 0; float solubility = 1.0f, equrate = 1.0f, friction = 1.0f;
  int erodes = 0; float erosionrate = 0.0f;
  int cascades = 0; float maxdiff = 1.0f, settling = 0.0f;
  int abrades = 0; float suspension = 0.0f, abrasion = 0.0f;
};
struct LayerEntry {           // SurfLayer, surface.h:65-101
  int type = 0;
  float min = 0.0f, bias = 0.0f, scale = 1.0f, octaves = 1.0f, lacunarity = 1.0f, gain = 0.0f, frequency = 1.0f;
};
struct WorldEntry {           // SoilMachine.cpp:9-14
  int sizex = 256, sizey = 256, scale = 80, nwind = 250, nwater = 250;
};
struct SoilFile {
  std::vector<SoilEntry> soils;
  std::map<std::string, int> soilmap;
  std::vector<LayerEntry> layers;
  WorldEntry world;
};

inline SoilEntry air_entry() {  // the built-in soil 0, surface.h:43-49
  SoilEntry a;
  a.name = "Air";
  a.porosity = 1.0f;
  a.color[0] = 0.0f; a.color[1] = 0.2f; a.color[2] = 0.4f; a.color[3] = 1.0f;
  a.solubility = a.equrate = a.friction = a.maxdiff = 0.0f;
  return a;
}

namespace soilfile_detail {

enum Block { NONE, SOIL, LAYER, WORLD };
enum Decode { F32, I32, SOILREF, HEXRGB };
struct Key { const char* word; Block block; Decode decode; size_t offset; };

#define SM_SOIL_KEY(word, decode, field) {word, SOIL, decode, offsetof(SoilEntry, field)}
#define SM_LAYER_KEY(word, field) {word, LAYER, F32, offsetof(LayerEntry, field)}
#define SM_WORLD_KEY(word, field) {word, WORLD, I32, offsetof(WorldEntry, field)}
inline const Key* keys(size_t& n) {
  static const Key table[] = {
      SM_SOIL_KEY("TRANSPORTS", SOILREF, transports), SM_SOIL_KEY("ERODES", SOILREF, erodes),
      SM_SOIL_KEY("CASCADES", SOILREF, cascades),     SM_SOIL_KEY("ABRADES", SOILREF, abrades),
      SM_SOIL_KEY("DENSITY", F32, density),           SM_SOIL_KEY("POROSITY", F32, porosity),
      SM_SOIL_KEY("COLOR", HEXRGB, color),            SM_SOIL_KEY("SOLUBILITY", F32, solubility),
      SM_SOIL_KEY("EQUILIBRIUM", F32, equrate),       SM_SOIL_KEY("FRICTION", F32, friction),
      SM_SOIL_KEY("EROSIONRATE", F32, erosionrate),   SM_SOIL_KEY("MAXDIFF", F32, maxdiff),
      SM_SOIL_KEY("SETTLING", F32, settling),         SM_SOIL_KEY("SUSPENSION", F32, suspension),
      SM_SOIL_KEY("ABRASION", F32, abrasion),
      {"Ka", SOIL, F32, offsetof(SoilEntry, phong) + 0},  {"Kd", SOIL, F32, offsetof(SoilEntry, phong) + 4},
      {"Ks", SOIL, F32, offsetof(SoilEntry, phong) + 8},  {"Kk", SOIL, F32, offsetof(SoilEntry, phong) + 12},
      SM_LAYER_KEY("MIN", min),         SM_LAYER_KEY("BIAS", bias),   SM_LAYER_KEY("SCALE", scale),
      SM_LAYER_KEY("OCTAVES", octaves), SM_LAYER_KEY("LACUNARITY", lacunarity), SM_LAYER_KEY("GAIN", gain),
      SM_LAYER_KEY("FREQUENCY", frequency),
      SM_WORLD_KEY("SIZEX", sizex), SM_WORLD_KEY("SIZEY", sizey), SM_WORLD_KEY("SCALE", scale),
      SM_WORLD_KEY("NWIND", nwind), SM_WORLD_KEY("NWATER", nwater),
  };
  n = sizeof(table) / sizeof(table[0]);
  return table;
}
#undef SM_SOIL_KEY
#undef SM_LAYER_KEY
#undef SM_WORLD_KEY

inline int hexdigit(char ch) {
  if (ch >= '0' && ch <= '9') return ch - '0';
  if (ch >= 'A' && ch <= 'F') return ch - 'A' + 10;
  return -1;
}

struct Reader {
  const std::string& file;
  SoilFile& out;
  SoilEntry running;          // the record every SOIL block edits; never reset between blocks
  Block block = NONE;
  bool open = false;
  int lineno = 0;

  [[noreturn]] void bad() const {
    throw SoilFileError("Incorrect Syntax in Line " + std::to_string(lineno) + " of " + file);
  }
  int soil_id(const std::string& name) {   // first mention allocates the id (with the running record as a placeholder)
    auto it = out.soilmap.find(name);
    if (it != out.soilmap.end()) return it->second;
    const int id = (int)out.soils.size();
    out.soilmap.emplace(name, id);
    out.soils.push_back(running);
    return id;
  }
  // "<name> {" -> name (the character before the brace is the separating blank)
  std::string header_name(const std::string& rest) const {
    const size_t brace = rest.find('{');
    if (brace == std::string::npos) bad();
    return rest.substr(0, brace - 1);
  }
  void store(void* base, const Key& k, const std::string& value) {
    char* dst = static_cast<char*>(base) + k.offset;
    switch (k.decode) {
      case F32: { const float v = std::stof(value); std::memcpy(dst, &v, sizeof v); break; }
      case I32: { const int v = std::stoi(value); std::memcpy(dst, &v, sizeof v); break; }
      case SOILREF: { const int v = soil_id(value); std::memcpy(dst, &v, sizeof v); break; }
      case HEXRGB: {
        if (value.size() < 6) bad();
        for (char ch : value) if (hexdigit(ch) < 0) bad();
        float rgba[4];
        for (int c = 0; c < 3; c++) rgba[c] = (float)(16 * hexdigit(value[2 * c]) + hexdigit(value[2 * c + 1])) / 255.0f;
        rgba[3] = 1.0f;
        std::memcpy(dst, rgba, sizeof rgba);
        break;
      }
    }
  }
  void line(std::string text) {
    lineno++;
    const size_t hash = text.find('#');
    if (hash != std::string::npos) text.erase(hash);
    if (text.empty()) return;
    if (text == "}") {
      if (!open) bad();
      if (block == SOIL) out.soils[(size_t)out.soilmap[running.name]] = running;
      open = false;
      return;
    }
    const size_t blank = text.find(' ');
    if (blank == std::string::npos) bad();
    const std::string word = text.substr(0, blank), rest = text.substr(blank + 1);
    if (word == "SOIL") {
      running.name = header_name(rest);
      soil_id(running.name);
      block = SOIL; open = true;
    } else if (word == "LAYER") {
      running.name = header_name(rest);
      auto it = out.soilmap.find(running.name);
      if (it == out.soilmap.end()) bad();              // a layer must name a soil that exists already
      LayerEntry l; l.type = it->second;
      out.layers.push_back(l);
      block = LAYER; open = true;
    } else if (word == "WORLD") {
      if (rest.find('{') == std::string::npos) bad();
      block = WORLD; open = true;
    } else {
      void* base = block == SOIL ? (void*)&running : block == LAYER ? (void*)&out.layers.back()
                 : block == WORLD ? (void*)&out.world : nullptr;
      if (!base) return;
      size_t n; const Key* k = keys(n);
      for (size_t i = 0; i < n; i++)
        if (k[i].block == block && word == k[i].word) { store(base, k[i], rest); break; }
    }
  }
};

}  // namespace soilfile_detail

// Reads `file` into `out`.  `out` may already hold tables (further files append, as the reference's globals
// do); pass a fresh SoilFile for the usual one-file case.
inline void parse_soil_file(const std::string& file, SoilFile& out) {
  std::ifstream in(file);
  if (!in.is_open()) throw SoilFileError("Failed to open soil profile " + file);
  if (out.soils.empty()) { out.soils.push_back(air_entry()); out.soilmap["Air"] = 0; }
  soilfile_detail::Reader rd{file, out};
  for (std::string text; std::getline(in, text);) rd.line(text);
}

}  // namespace soilmachine
