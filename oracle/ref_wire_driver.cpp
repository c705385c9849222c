// oracle/ref_wire_driver.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// Runs the reference's own, unmodified SerialPort::getDataFromSerial (src/SerialPort.cpp:106-138) and the formatting of
// SerialPort::writePort (cpp:165) on sample frames and prints JSON goldens for tests/golden/wire_ref.json.  The SerialPort
// object is NOT constructed (its constructor loops until /dev/ttyUSB0 opens, cpp:37-50); getDataFromSerial touches no member.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "SerialPort.h"

int main() {
  SerialPort *sp = static_cast<SerialPort *>(std::calloc(1, sizeof(SerialPort)));
  const std::vector<std::string> frames = {
      "0.0150 0.0100 0.0000 0.0200 0.0000 \r\n",
      "0.0149 -0.1234567 1.5000001 -0.0499999 12.3456789\n",
      "0.02 1e-3 -2.5e-2 3.25 -4.125 extra 9 9\n",
      "0.0151 0.33333333333 0.66666666667 -0.11111111111 0.999999999\n",
      "15 100 -200 300.5 -400.25                \n",
      "0.0150 0.0100 abc 0.0200 0.0000 padpadpadpad\n",
  };
  std::printf("{\"frames\": [\n");
  for (size_t f = 0; f < frames.size(); ++f) {
    char buf[64];
    std::memset(buf, 0, sizeof(buf));
    std::strncpy(buf, frames[f].c_str(), sizeof(buf) - 1);
    double dt = -1.0;
    Eigen::Matrix<double, N_S, 1> X;
    X.setConstant(-1.0);
    sp->getDataFromSerial(dt, X, buf);
    std::string esc;
    for (char c : frames[f]) { if (c == '\n') esc += "\\n"; else if (c == '\r') esc += "\\r"; else esc += c; }
    std::printf("  {\"text\": \"%s\", \"dt\": %.17g, \"X\": [%.17g, %.17g, %.17g, %.17g]}%s\n", esc.c_str(), dt, X(0), X(1), X(2), X(3),
                f + 1 < frames.size() ? "," : "");
  }
  std::printf("],\n\"controls\": [\n");
  const std::vector<double> us = {0.0, 1.0, -0.4, 123.456789, -255.0, 0.000012345, 1e6, -3.14159265358979};
  for (size_t k = 0; k < us.size(); ++k) {
    const std::string s = std::to_string(us[k]);                 // what writePort formats (cpp:165)
    const size_t sent = sizeof(s.c_str());                       // ... and how many bytes it hands to write(): sizeof(char*) = 8
    std::printf("  {\"U\": %.17g, \"text\": \"%s\", \"bytes_sent\": %zu}%s\n", us[k], s.c_str(), sent, k + 1 < us.size() ? "," : "");
  }
  std::printf("]}\n");
  return 0;
}
