// TEST INFRASTRUCTURE ONLY -- shim for "rapidxml-1.13/rapidxml_utils.hpp" (core.h:9): rapidxml::file<>, which
// loads a whole file into a zero-terminated buffer (model.cpp:226).  Like the original it pulls in <fstream>,
// which core.cpp:48 relies on.  See rapidxml.hpp in this directory.
#ifndef ORACLE_SHIM_RAPIDXML_UTILS_HPP
#define ORACLE_SHIM_RAPIDXML_UTILS_HPP
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>
#include "rapidxml.hpp"

namespace rapidxml {
template <class Ch = char>
class file {
 public:
  explicit file(const char* filename) {
    std::basic_ifstream<Ch> stream(filename, std::ios::binary);
    if (!stream) throw std::runtime_error(std::string("cannot open file ") + filename);
    stream.unsetf(std::ios::skipws);
    stream.seekg(0, std::ios::end);
    size_t size = (size_t)stream.tellg();
    stream.seekg(0);
    m_data.resize(size + 1);
    stream.read(&m_data.front(), static_cast<std::streamsize>(size));
    m_data[size] = 0;
  }
  Ch* data() { return &m_data.front(); }
  const Ch* data() const { return &m_data.front(); }
  std::size_t size() const { return m_data.size(); }

 private:
  std::vector<Ch> m_data;
};
}  // namespace rapidxml
#endif
