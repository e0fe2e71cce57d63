// Minimal YAML reader for the reference's configuration files (material.yaml, numerics.yaml, load.yaml): block and flow
// mappings and sequences, plain / quoted scalars, comments, document markers -- the subset those files use.  The reference
// parses them with libfyaml on the Fortran side (/root/reference/src/YAML_parse.f90, YAML_types.f90) and with libyaml in
// the C++ driver (/root/reference/src/grid/config.cpp:247-463); neither library is a dependency of this one.
#pragma once
#include <memory>
#include <string>
#include <utility>
#include <vector>

struct YamlNode {
  enum Kind { Null, Scalar, Map, Seq } kind = Null;
  std::string scalar;
  std::vector<std::pair<std::string, std::shared_ptr<YamlNode>>> map;
  std::vector<std::shared_ptr<YamlNode>> seq;

  bool is_map() const { return kind == Map; }
  bool is_seq() const { return kind == Seq; }
  bool is_scalar() const { return kind == Scalar; }
  const YamlNode* get(const std::string& key) const;            // nullptr if absent or not a map
  bool has(const std::string& key) const { return get(key) != nullptr; }
  double as_double() const;                                     // throws std::runtime_error on a non-number
  long as_int() const;
  bool as_bool() const;
  std::vector<double> as_doubles() const;                       // a sequence of numbers, or a single number
  std::vector<int> as_ints() const;
};

// throws std::runtime_error with a line number on input outside the subset
std::shared_ptr<YamlNode> yaml_parse(const std::string& text);
std::shared_ptr<YamlNode> yaml_parse_file(const std::string& path);
