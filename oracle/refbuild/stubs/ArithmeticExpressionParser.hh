// stand-in for the bison-generated header of src/Core/ArithmeticExpressionParser.yy (no bison in this image)
#pragma once
#include <string>
namespace Core {
class ArithmeticExpressionParserDriver {
public:
    bool parse(const std::string&, double& result) {
        result = 0;
        return false;
    }
    std::string getLastError() const {
        return ": arithmetic expressions are not available in the oracle build";
    }
};
}  // namespace Core
