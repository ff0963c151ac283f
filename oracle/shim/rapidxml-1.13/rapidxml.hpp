// TEST INFRASTRUCTURE ONLY -- shim for "rapidxml-1.13/rapidxml.hpp" (core.h:8), a third-party header-only XML
// parser that is absent from /root/reference and from this image.  It offers the small part of rapidxml's DOM
// interface the reference calls (model.cpp:226-275, visualization.cpp:442-480, core.cpp:14-28): xml_document<>::parse<0>,
// first_node(name), next_sibling(name), first_attribute(name), name(), value().  Elements, attributes (either
// quote), self-closing tags, comments, declarations and DOCTYPE are understood; character data is skipped.
#ifndef ORACLE_SHIM_RAPIDXML_HPP
#define ORACLE_SHIM_RAPIDXML_HPP
#include <cassert>  // the original includes it too; pergen.cpp:459 relies on that
#include <cstdlib>
#include <cstring>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

namespace rapidxml {

class parse_error : public std::runtime_error {
 public:
  explicit parse_error(const char* what) : std::runtime_error(what) {}
};

template <class Ch = char>
class xml_base {
 public:
  Ch* name() const { return const_cast<Ch*>(name_.c_str()); }
  Ch* value() const { return const_cast<Ch*>(value_.c_str()); }
  std::basic_string<Ch> name_, value_;
};

template <class Ch = char>
class xml_attribute : public xml_base<Ch> {
 public:
  xml_attribute() : next_(0) {}
  xml_attribute<Ch>* next_attribute(const Ch* name = 0) const {
    for (xml_attribute<Ch>* a = next_; a; a = a->next_)
      if (!name || a->name_ == name) return a;
    return 0;
  }
  xml_attribute<Ch>* next_;
};

template <class Ch = char>
class xml_node : public xml_base<Ch> {
 public:
  xml_node() : parent_(0), first_child_(0), last_child_(0), next_(0), first_attr_(0), last_attr_(0) {}
  virtual ~xml_node() {}
  xml_node<Ch>* parent() const { return parent_; }
  xml_node<Ch>* first_node(const Ch* name = 0) const {
    for (xml_node<Ch>* c = first_child_; c; c = c->next_)
      if (!name || c->name_ == name) return c;
    return 0;
  }
  xml_node<Ch>* next_sibling(const Ch* name = 0) const {
    for (xml_node<Ch>* c = next_; c; c = c->next_)
      if (!name || c->name_ == name) return c;
    return 0;
  }
  xml_attribute<Ch>* first_attribute(const Ch* name = 0) const {
    for (xml_attribute<Ch>* a = first_attr_; a; a = a->next_)
      if (!name || a->name_ == name) return a;
    return 0;
  }
  void append_node(xml_node<Ch>* c) {
    c->parent_ = this;
    if (last_child_) last_child_->next_ = c; else first_child_ = c;
    last_child_ = c;
  }
  void append_attribute(xml_attribute<Ch>* a) {
    if (last_attr_) last_attr_->next_ = a; else first_attr_ = a;
    last_attr_ = a;
  }
  xml_node<Ch>*parent_, *first_child_, *last_child_, *next_;
  xml_attribute<Ch>*first_attr_, *last_attr_;
};

template <class Ch = char>
class xml_document : public xml_node<Ch> {
 public:
  xml_document() {}
  ~xml_document() { clear(); }
  void clear() {
    for (size_t i = 0; i < nodes_.size(); i++) delete nodes_[i];
    for (size_t i = 0; i < attrs_.size(); i++) delete attrs_[i];
    nodes_.clear();
    attrs_.clear();
    this->first_child_ = this->last_child_ = 0;
  }
  template <int Flags>
  void parse(Ch* text) {
    clear();
    const Ch* p = text;
    xml_node<Ch>* cur = this;
    while (*p) {
      if (*p != '<') { p++; continue; }  // character data
      if (!std::strncmp(p, "<!--", 4)) { p = find(p + 4, "-->") + 3; continue; }
      if (p[1] == '?') { p = find(p + 2, "?>") + 2; continue; }
      if (p[1] == '!') { p = find(p + 2, ">") + 1; continue; }
      if (p[1] == '/') {
        p = find(p + 2, ">") + 1;
        if (cur == this) throw parse_error("unbalanced closing tag");
        cur = cur->parent();
        continue;
      }
      p++;
      xml_node<Ch>* node = new xml_node<Ch>;
      nodes_.push_back(node);
      const Ch* s = p;
      while (*p && !is_space(*p) && *p != '>' && *p != '/') p++;
      node->name_.assign(s, p);
      cur->append_node(node);
      for (;;) {
        while (is_space(*p)) p++;
        if (!*p) throw parse_error("unexpected end of data");
        if (*p == '/') { p = find(p, ">") + 1; break; }        // <name ... />
        if (*p == '>') { p++; cur = node; break; }
        xml_attribute<Ch>* a = new xml_attribute<Ch>;
        attrs_.push_back(a);
        s = p;
        while (*p && !is_space(*p) && *p != '=') p++;
        a->name_.assign(s, p);
        while (is_space(*p)) p++;
        if (*p != '=') throw parse_error("expected =");
        p++;
        while (is_space(*p)) p++;
        Ch q = *p;
        if (q != '"' && q != '\'') throw parse_error("expected quote");
        s = ++p;
        while (*p && *p != q) p++;
        if (!*p) throw parse_error("unterminated attribute value");
        a->value_.assign(s, p);
        p++;
        node->append_attribute(a);
      }
    }
  }

 private:
  static bool is_space(Ch c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r'; }
  static const Ch* find(const Ch* p, const char* pat) {
    const Ch* r = std::strstr(p, pat);
    if (!r) throw parse_error("unexpected end of data");
    return r;
  }
  std::vector<xml_node<Ch>*> nodes_;
  std::vector<xml_attribute<Ch>*> attrs_;
};

}  // namespace rapidxml
#endif
