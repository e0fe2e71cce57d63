// Test helper: parses a YAML file with the library's minimal reader (csrc/yaml_min.cpp) and prints the tree as JSON with every
// scalar as a string, for comparison with PyYAML (tests/test_yaml_cpu.py).
#include <cstdio>
#include <string>
#include "../../solver_cpp_b200/csrc/yaml_min.h"

static void esc(const std::string& s) {
  putchar('"');
  for (char ch : s) { if (ch == '"' || ch == '\\') putchar('\\'); if (ch == '\n') { fputs("\\n", stdout); continue; } putchar(ch); }
  putchar('"');
}
static void dump(const YamlNode& n) {
  switch (n.kind) {
    case YamlNode::Null: fputs("null", stdout); break;
    case YamlNode::Scalar: esc(n.scalar); break;
    case YamlNode::Seq: putchar('['); for (size_t i = 0; i < n.seq.size(); i++) { if (i) putchar(','); dump(*n.seq[i]); } putchar(']'); break;
    case YamlNode::Map: putchar('{'); for (size_t i = 0; i < n.map.size(); i++) { if (i) putchar(','); esc(n.map[i].first); putchar(':'); dump(*n.map[i].second); } putchar('}'); break;
  }
}
int main(int argc, char** argv) {
  if (argc < 2) return 2;
  try { dump(*yaml_parse_file(argv[1])); putchar('\n'); }
  catch (const std::exception& e) { fprintf(stderr, "%s\n", e.what()); return 1; }
  return 0;
}
