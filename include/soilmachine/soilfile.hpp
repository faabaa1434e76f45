// include/soilmachine/soilfile.hpp -- the `.soil` text format, restated from the reference loader
// loadsoil() (source/io.h:7-230) including its quirks, because presets written for the reference must
// produce the same tables here:
//   * `#` starts a comment; only exactly-empty lines are skipped (io.h:43-47);
//   * ONE SurfParam object is reused for all SOIL blocks and never reset, so a block inherits every field
//     it does not set from the previous block (io.h:35);
//   * soil ids are assigned in order of first mention - `TRANSPORTS X` before `SOIL X` pushes a placeholder
//     copy of the current parameters (io.h:125-152); "Air" is id 0 (surface.h:41-57);
//   * `}` on its own line stores the block (io.h:50-58); LAYER order = deposition order (io.h:88-109);
//   * WORLD accepts only SIZEX SIZEY SCALE NWIND NWATER (io.h:208-220; a SEED line is ignored);
//   * colours are six upper-case hex digits -> rgb/255, alpha 1 (io.h:23-33).
// Errors throw SoilFileError instead of the reference's `cout + exit(0)`.
#pragma once
#include <fstream>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace soilmachine {

struct SoilFileError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

struct SoilEntry {            // SurfParam, surface.h:11-39
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

inline SoilEntry air_entry() {  // surface.h:43-49
  SoilEntry a;
  a.name = "Air"; a.density = 0.0f; a.porosity = 1.0f;
  a.color[0] = 0.0f; a.color[1] = 0.2f; a.color[2] = 0.4f; a.color[3] = 1.0f;
  a.transports = 0; a.solubility = 0.0f; a.equrate = 0.0f; a.friction = 0.0f;
  a.erodes = 0; a.erosionrate = 0.0f; a.cascades = 0; a.maxdiff = 0.0f; a.settling = 0.0f;
  a.abrades = 0; a.suspension = 0.0f; a.abrasion = 0.0f;
  return a;
}

// io.h:7-230.  `out` may already hold tables (the reference appends to its globals); pass a fresh SoilFile
// for the usual one-file case.
inline void parse_soil_file(const std::string& file, SoilFile& out) {
  std::ifstream in(file, std::ios::in);
  if (!in.is_open()) throw SoilFileError("Failed to open soil profile " + file);
  if (out.soils.empty()) { out.soils.push_back(air_entry()); out.soilmap["Air"] = 0; }
  std::string line;
  int linenr = 0;
  auto syntaxerr = [&]() { throw SoilFileError("Incorrect Syntax in Line " + std::to_string(linenr) + " of " + file); };
  auto hexcol = [&](const std::string& h, float* c) {
    if (h.size() < 6) syntaxerr();
    const std::string allowed = "0123456789ABCDEF";
    for (char ch : h) if (allowed.find(ch) == std::string::npos) syntaxerr();
    const float R = 16 * allowed.find(h[0]) + allowed.find(h[1]);
    const float G = 16 * allowed.find(h[2]) + allowed.find(h[3]);
    const float B = 16 * allowed.find(h[4]) + allowed.find(h[5]);
    c[0] = R / 255.0f; c[1] = G / 255.0f; c[2] = B / 255.0f; c[3] = (float)255.0 / 255.0f;
  };
  auto mention = [&](const std::string& name, const SoilEntry& param) {   // id by first mention
    if (!out.soilmap.count(name)) { out.soilmap[name] = (int)out.soils.size(); out.soils.push_back(param); }
    return out.soilmap[name];
  };
  SoilEntry param;            // reused, never reset
  bool open = false;
  std::string soillayer;
  while (std::getline(in, line)) {
    linenr++;
    size_t found = line.find('#');
    if (found != std::string::npos) line = line.substr(0, found);
    if (line == "") continue;
    if (line == "}") {
      if (!open) syntaxerr();
      if (soillayer == "SOIL") out.soils[out.soilmap[param.name]] = param;
      open = false;
      continue;
    }
    found = line.find(' ');
    if (found == std::string::npos) syntaxerr();
    const std::string tag = line.substr(0, found);
    const std::string val = line.substr(found + 1);
    if (tag == "SOIL") {
      found = val.find('{');
      if (found == std::string::npos) syntaxerr();
      param.name = val.substr(0, found - 1);
      mention(param.name, param);
      soillayer = tag; open = true;
      continue;
    }
    if (tag == "LAYER") {
      found = val.find('{');
      if (found == std::string::npos) syntaxerr();
      param.name = val.substr(0, found - 1);
      if (!out.soilmap.count(param.name)) syntaxerr();
      LayerEntry l; l.type = out.soilmap[param.name];
      out.layers.push_back(l);
      soillayer = tag; open = true;
      continue;
    }
    if (tag == "WORLD") {
      if (val.find('{') == std::string::npos) syntaxerr();
      soillayer = tag; open = true;
      continue;
    }
    if (soillayer == "SOIL") {
      if (tag == "TRANSPORTS") param.transports = mention(val, param);
      if (tag == "ERODES") param.erodes = mention(val, param);
      if (tag == "CASCADES") param.cascades = mention(val, param);
      if (tag == "ABRADES") param.abrades = mention(val, param);
      if (tag == "DENSITY") param.density = std::stof(val);
      if (tag == "POROSITY") param.porosity = std::stof(val);
      if (tag == "COLOR") hexcol(val, param.color);
      if (tag == "SOLUBILITY") param.solubility = std::stof(val);
      if (tag == "EQUILIBRIUM") param.equrate = std::stof(val);
      if (tag == "FRICTION") param.friction = std::stof(val);
      if (tag == "EROSIONRATE") param.erosionrate = std::stof(val);
      if (tag == "MAXDIFF") param.maxdiff = std::stof(val);
      if (tag == "SETTLING") param.settling = std::stof(val);
      if (tag == "SUSPENSION") param.suspension = std::stof(val);
      if (tag == "ABRASION") param.abrasion = std::stof(val);
      if (tag == "Ka") param.phong[0] = std::stof(val);
      if (tag == "Kd") param.phong[1] = std::stof(val);
      if (tag == "Ks") param.phong[2] = std::stof(val);
      if (tag == "Kk") param.phong[3] = std::stof(val);
    }
    if (soillayer == "LAYER") {
      LayerEntry& l = out.layers.back();
      if (tag == "MIN") l.min = std::stof(val);
      if (tag == "BIAS") l.bias = std::stof(val);
      if (tag == "SCALE") l.scale = std::stof(val);
      if (tag == "OCTAVES") l.octaves = std::stof(val);
      if (tag == "LACUNARITY") l.lacunarity = std::stof(val);
      if (tag == "GAIN") l.gain = std::stof(val);
      if (tag == "FREQUENCY") l.frequency = std::stof(val);
    }
    if (soillayer == "WORLD") {
      if (tag == "SIZEX") out.world.sizex = std::stoi(val);
      if (tag == "SIZEY") out.world.sizey = std::stoi(val);
      if (tag == "SCALE") out.world.scale = std::stoi(val);
      if (tag == "NWIND") out.world.nwind = std::stoi(val);
      if (tag == "NWATER") out.world.nwater = std::stoi(val);
    }
  }
}

}  // namespace soilmachine
