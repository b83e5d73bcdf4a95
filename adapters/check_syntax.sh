#!/bin/bash
# Header check of the adapters against the real RASR headers (g++ -fsyntax-only).  The reference cannot be built
# here (no libxml2 / boost), so two tiny stand-in headers under adapters/stubs/ replace the only third-party
# includes the RASR headers pull in (libxml/parser.h: typedefs used by Core/XmlParser.hh;
# boost/thread/synchronized_value.hpp: used by Core/Configuration.hh).  Usage: check_syntax.sh [reference-root]
REF=${1:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
rc=0
for f in B200FeatureScorer.cc B200MfccNode.cc B200NnNetwork.cc B200Nodes.cc B200LinearSearch.cc Module.cc; do
  g++ -fsyntax-only -std=gnu++20 -funsigned-char -I"$HERE/stubs" -I"$REF/src" -I"$HERE/../include" "$HERE/$f" || rc=1
done
exit $rc
