"oracle build of the reference sources (no generated SourceVersion.cc)\n"
