/* minixml: the slice of libxml2's SAX1 interface that src/Core/XmlParser.{hh,cc} uses.  Stand-in written for the
 * oracle build (no libxml2 in this image); implementation in oracle/refbuild/minixml.cc. */
#ifndef MINIXML_PARSER_H
#define MINIXML_PARSER_H
#include <stdarg.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef unsigned char xmlChar;
typedef struct _xmlParserInput*   xmlParserInputPtr;
typedef struct _xmlEntity*        xmlEntityPtr;
typedef struct _xmlEnumeration*   xmlEnumerationPtr;
typedef struct _xmlElementContent* xmlElementContentPtr;
typedef struct _xmlDict*          xmlDictPtr;

/* callback table, libxml2's xmlSAXHandler field order (SAX1 part) */
typedef struct _xmlSAXHandler {
    void (*internalSubset)(void*, const xmlChar*, const xmlChar*, const xmlChar*);
    int (*isStandalone)(void*);
    int (*hasInternalSubset)(void*);
    int (*hasExternalSubset)(void*);
    xmlParserInputPtr (*resolveEntity)(void*, const xmlChar*, const xmlChar*);
    xmlEntityPtr (*getEntity)(void*, const xmlChar*);
    void (*entityDecl)(void*, const xmlChar*, int, const xmlChar*, const xmlChar*, xmlChar*);
    void (*notationDecl)(void*, const xmlChar*, const xmlChar*, const xmlChar*);
    void (*attributeDecl)(void*, const xmlChar*, const xmlChar*, int, int, const xmlChar*, xmlEnumerationPtr);
    void (*elementDecl)(void*, const xmlChar*, int, xmlElementContentPtr);
    void (*unparsedEntityDecl)(void*, const xmlChar*, const xmlChar*, const xmlChar*, const xmlChar*);
    void (*setDocumentLocator)(void*, void*);
    void (*startDocument)(void*);
    void (*endDocument)(void*);
    void (*startElement)(void*, const xmlChar*, const xmlChar**);
    void (*endElement)(void*, const xmlChar*);
    void (*reference)(void*, const xmlChar*);
    void (*characters)(void*, const xmlChar*, int);
    void (*ignorableWhitespace)(void*, const xmlChar*, int);
    void (*processingInstruction)(void*, const xmlChar*, const xmlChar*);
    void (*comment)(void*, const xmlChar*);
    void (*warning)(void*, const char*, ...);
    void (*error)(void*, const char*, ...);
    void (*fatalError)(void*, const char*, ...);
    xmlEntityPtr (*getParameterEntity)(void*, const xmlChar*);
    void (*cdataBlock)(void*, const xmlChar*, int);
    void (*externalSubset)(void*, const xmlChar*, const xmlChar*, const xmlChar*);
    unsigned int initialized;
} xmlSAXHandler;

typedef struct _xmlParserCtxt {
    xmlSAXHandler* sax;
    void*          userData;
    int            options;
    int            wellFormed;
    int            valid;
    int            errNo;
    void*          priv; /* minixml state */
} xmlParserCtxt;
typedef xmlParserCtxt* xmlParserCtxtPtr;

#define XML_PARSE_HUGE (1 << 19)

xmlParserCtxtPtr xmlCreateMemoryParserCtxt(const char* buffer, int size);
xmlParserCtxtPtr xmlCreateFileParserCtxt(const char* filename);
xmlParserCtxtPtr xmlCreatePushParserCtxt(xmlSAXHandler* sax, void* userData, const char* chunk, int size,
                                         const char* filename);
int  xmlParseChunk(xmlParserCtxtPtr ctxt, const char* chunk, int size, int terminate);
int  xmlParseDocument(xmlParserCtxtPtr ctxt);
void xmlFreeParserCtxt(xmlParserCtxtPtr ctxt);
xmlEntityPtr xmlGetPredefinedEntity(const xmlChar* name);
int  xmlStrcmp(const xmlChar* a, const xmlChar* b);
const xmlChar* xmlSAX2GetSystemId(void* ctx);
int  xmlSAX2GetLineNumber(void* ctx);
int  xmlSAX2GetColumnNumber(void* ctx);
xmlParserInputPtr resolveEntity(void* ctx, const xmlChar* publicId, const xmlChar* systemId);
/* src/Math/Random.cc:31-37 creates and frees a dictionary once to initialise libxml2's random seed */
xmlDictPtr xmlDictCreate(void);
void       xmlDictFree(xmlDictPtr dict);
#ifdef __cplusplus
}
#endif
#endif
