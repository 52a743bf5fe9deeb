// Minimal XML reader for the MJCF subset (host only). Replaces the parse half of
// mj_parseXMLString / mj_loadXML as called at reference src/physics.rs:13,19.
#pragma once
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace ox {

struct XmlError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

struct XmlElem {
  std::string name;
  int line = 0;
  std::vector<std::pair<std::string, std::string>> attrs;  // document order
  std::vector<std::unique_ptr<XmlElem>> children;

  const std::string* attr(const std::string& k) const {
    for (auto& a : attrs)
      if (a.first == k) return &a.second;
    return nullptr;
  }
};

class XmlParser {
 public:
  explicit XmlParser(const std::string& text) : s_(text) {}

  std::unique_ptr<XmlElem> parse() {
    skip_misc();
    if (eof()) fail("empty document");
    auto root = element();
    skip_misc();
    if (!eof()) fail("content after the root element");
    return root;
  }

 private:
  const std::string& s_;
  size_t p_ = 0;
  int line_ = 1;

  bool eof() const { return p_ >= s_.size(); }
  char cur() const { return s_[p_]; }
  void adv() {
    if (s_[p_] == '\n') ++line_;
    ++p_;
  }
  bool starts(const char* lit) const { return s_.compare(p_, std::char_traits<char>::length(lit), lit) == 0; }
  [[noreturn]] void fail(const std::string& msg) const {
    throw XmlError("XML parse error at line " + std::to_string(line_) + ": " + msg);
  }
  void skip_ws() {
    while (!eof() && (cur() == ' ' || cur() == '\t' || cur() == '\n' || cur() == '\r')) adv();
  }
  void skip_until(const char* lit) {
    while (!eof() && !starts(lit)) adv();
    if (eof()) fail(std::string("unterminated construct, expected '") + lit + "'");
    for (const char* c = lit; *c; ++c) adv();
  }
  // whitespace, comments, processing instructions, doctype
  void skip_misc() {
    for (;;) {
      skip_ws();
      if (eof()) return;
      if (starts("<!--")) skip_until("-->");
      else if (starts("<?")) skip_until("?>");
      else if (starts("<!")) skip_until(">");
      else return;
    }
  }
  static bool name_char(char c) {
    return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || (c >= '0' && c <= '9') || c == '_' || c == '-' || c == ':' || c == '.';
  }
  std::string name() {
    size_t b = p_;
    while (!eof() && name_char(cur())) adv();
    if (p_ == b) fail("expected a name");
    return s_.substr(b, p_ - b);
  }
  static std::string unescape(const std::string& v) {
    if (v.find('&') == std::string::npos) return v;
    static const std::pair<const char*, char> ents[] = {{"&lt;", '<'}, {"&gt;", '>'}, {"&amp;", '&'}, {"&quot;", '"'}, {"&apos;", '\''}};
    std::string o;
    for (size_t i = 0; i < v.size();) {
      bool hit = false;
      if (v[i] == '&')
        for (auto& e : ents) {
          size_t n = std::char_traits<char>::length(e.first);
          if (v.compare(i, n, e.first) == 0) {
            o += e.second;
            i += n;
            hit = true;
            break;
          }
        }
      if (!hit) o += v[i++];
    }
    return o;
  }
  std::unique_ptr<XmlElem> element() {
    if (eof() || cur() != '<') fail("expected '<'");
    adv();
    auto e = std::make_unique<XmlElem>();
    e->line = line_;
    e->name = name();
    for (;;) {
      skip_ws();
      if (eof()) fail("unterminated tag <" + e->name + ">");
      if (cur() == '/') {
        adv();
        if (eof() || cur() != '>') fail("expected '>' after '/'");
        adv();
        return e;
      }
      if (cur() == '>') {
        adv();
        break;
      }
      std::string k = name();
      skip_ws();
      if (eof() || cur() != '=') fail("expected '=' after attribute '" + k + "'");
      adv();
      skip_ws();
      if (eof() || (cur() != '"' && cur() != '\'')) fail("attribute value must be quoted");
      char q = cur();
      adv();
      size_t b = p_;
      while (!eof() && cur() != q) adv();
      if (eof()) fail("unterminated attribute value");
      std::string v = s_.substr(b, p_ - b);
      adv();
      if (e->attr(k)) fail("duplicate attribute '" + k + "'");
      e->attrs.emplace_back(k, unescape(v));
    }
    // content
    for (;;) {
      // text is ignored in MJCF
      while (!eof() && cur() != '<') adv();
      if (eof()) fail("missing </" + e->name + ">");
      if (starts("<!--")) {
        skip_until("-->");
        continue;
      }
      if (starts("<?")) {
        skip_until("?>");
        continue;
      }
      if (starts("</")) {
        adv();
        adv();
        std::string n = name();
        if (n != e->name) fail("mismatched closing tag </" + n + "> for <" + e->name + ">");
        skip_ws();
        if (eof() || cur() != '>') fail("expected '>'");
        adv();
        return e;
      }
      e->children.push_back(element());
    }
  }
};

}  // namespace ox
