#pragma once
typedef unsigned char xmlChar;
typedef struct _xmlParserInput* xmlParserInputPtr;
typedef struct _xmlEntity* xmlEntityPtr;
typedef struct _xmlEnumeration* xmlEnumerationPtr;
typedef struct _xmlElementContent* xmlElementContentPtr;
typedef struct _xmlParserCtxt* xmlParserCtxtPtr;
typedef struct _xmlSAXHandler { int dummy; } xmlSAXHandler;
