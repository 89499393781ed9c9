#include "xsbh_units.h"

#include <cctype>
#include <cmath>
#include <cstdlib>
#include <map>

namespace xsbh {
namespace {

// factor to internal units of every base / derived unit name
const std::map<std::string, double>& table() {
  static const std::map<std::string, double> t = {
      // length (internal: angstrom)
      {"m", 1e10}, {"meter", 1e10}, {"cm", 1e8}, {"mm", 1e7}, {"um", 1e4}, {"nm", 10.0}, {"ang", 1.0}, {"angstrom", 1.0},
      // mass (internal: Da)
      {"kg", 1.0 / kDalton}, {"g", 1e-3 / kDalton}, {"Da", 1.0}, {"Dalton", 1.0}, {"amu", 1.0},
      // time (internal: ps)
      {"s", 1e12}, {"second", 1e12}, {"ms", 1e9}, {"us", 1e6}, {"ns", 1e3}, {"ps", 1.0}, {"picosecond", 1.0}, {"fs", 1e-3},
      // charge (internal: elementary charge)
      {"C", 1.0 / kElementaryCharge}, {"e-", 1.0}, {"e", 1.0},
      // temperature, amount, luminosity, angle
      {"K", 1.0}, {"kelvin", 1.0}, {"mol", kAvogadro}, {"particle", 1.0}, {"cd", 1.0}, {"rad", 1.0}, {"radian", 1.0},
      {"degree", M_PI / 180.0}, {"deg", M_PI / 180.0},
      // energy and other derived units
      {"J", 1.0 / kInternalEnergyJ}, {"joule", 1.0 / kInternalEnergyJ}, {"eV", kEv},
      {"cal", 4.184 / kInternalEnergyJ}, {"kcal", 4184.0 / kInternalEnergyJ},
      {"N", 1.0 / kInternalEnergyJ * 1e-10}, {"Pa", 1.0 / kInternalEnergyJ * 1e-30}, {"bar", 1e5 / kInternalEnergyJ * 1e-30},
      {"GPa", 1e9 / kInternalEnergyJ * 1e-30}, {"atm", 101325.0 / kInternalEnergyJ * 1e-30},
      {"1", 1.0}};
  return t;
}

double one_unit(const std::string& tok) {
  // name[^power]
  std::string name = tok; int power = 1;
  size_t c = tok.find('^');
  if (c != std::string::npos) {
    name = tok.substr(0, c);
    char* e = nullptr;
    power = (int)std::strtol(tok.c_str() + c + 1, &e, 10);
    if (e == tok.c_str() + c + 1 || *e) throw UnitError("bad exponent in unit '" + tok + "'");
  }
  auto it = table().find(name);
  if (it == table().end()) throw UnitError("unknown unit '" + name + "'");
  return std::pow(it->second, power);
}

}  // namespace

double unit_factor(const std::string& expr) {
  double f = 1.0;
  char op = '*';
  std::string tok;
  auto flush = [&]() {
    if (tok.empty()) throw UnitError("malformed unit expression '" + expr + "'");
    double u = one_unit(tok);
    f = op == '/' ? f / u : f * u;
    tok.clear();
  };
  for (size_t i = 0; i < expr.size(); ++i) {
    char ch = expr[i];
    if (std::isspace((unsigned char)ch)) continue;
    if (ch == '*' || ch == '.' || ch == '/') {
      // "e-" is a unit name and '^-1' an exponent: '-' never separates; '.' separates only between names
      flush(); op = ch == '/' ? '/' : '*';
    } else tok += ch;
  }
  flush();
  return f;
}

double quantity(const std::string& text) {
  const char* b = text.c_str();
  char* e = nullptr;
  double v = std::strtod(b, &e);
  if (e == b) throw UnitError("number expected in quantity '" + text + "'");
  std::string unit(e);
  size_t i = 0; while (i < unit.size() && std::isspace((unsigned char)unit[i])) ++i;
  unit = unit.substr(i);
  while (!unit.empty() && std::isspace((unsigned char)unit.back())) unit.pop_back();
  if (unit.empty()) return v;
  return v * unit_factor(unit);
}

double quantity(const Node& n) {
  if (n.is_map()) {   // onika also accepts { value: x, unity: u }-style maps in old decks
    const Node* v = n.find("value"); const Node* u = n.find("unity");
    if (v && u) return v->as_double() * unit_factor(u->as_string());
  }
  return quantity(n.as_string());
}

}  // namespace xsbh
