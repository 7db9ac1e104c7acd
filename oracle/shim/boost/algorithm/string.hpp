// Stub: boost::split / is_any_of / token_compress_on as TrajectoryPlanner::reconfigure uses them (y_vels parsing).
#pragma once
#include <string>
#include <vector>
namespace boost {
struct is_any_of {
  std::string set;
  explicit is_any_of(const char* s) : set(s) {}
  bool operator()(char c) const { return set.find(c) != std::string::npos; }
};
enum token_compress_mode_type { token_compress_on, token_compress_off };
template <class Pred>
void split(std::vector<std::string>& out, const std::string& in, Pred pred, token_compress_mode_type mode = token_compress_off) {
  out.clear();
  std::string cur;
  bool last_was_sep = false;
  for (char c : in) {
    if (pred(c)) {
      if (!(mode == token_compress_on && last_was_sep)) { out.push_back(cur); cur.clear(); }
      last_was_sep = true;
    } else {
      cur.push_back(c);
      last_was_sep = false;
    }
  }
  out.push_back(cur);
}
}  // namespace boost
