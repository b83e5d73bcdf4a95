/*
 * minixml.cc -- a small non-validating SAX parser behind the libxml2 entry points that the reference's
 * src/Core/XmlParser.cc calls (see stubs/libxml/parser.h).
 *
 * TEST INFRASTRUCTURE ONLY (oracle/_ref build): this image has no libxml2, and without an XML parser the
 * reference's Flow network files (src/Tools/FeatureExtraction/share/*.flow) could not be read by the
 * reference's own Flow::NetworkParser.  Contains no reference code.  Supported: prolog / processing instructions,
 * comments, DOCTYPE (skipped), elements with attributes, character data, CDATA sections, the five predefined
 * entities and numeric character references.  Push mode collects the chunks and parses at the terminating call.
 */
#include <libxml/parser.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {
struct State {
    std::string text;
    std::string systemId;
    size_t      pos  = 0;
    int         line = 1, col = 1;
};

State* st(xmlParserCtxtPtr c) {
    return static_cast<State*>(c->priv);
}

xmlParserCtxtPtr make_ctxt(const char* data, size_t n, const char* name) {
    xmlParserCtxtPtr c = static_cast<xmlParserCtxtPtr>(calloc(1, sizeof(xmlParserCtxt)));
    State*           s = new State();
    if (data)
        s->text.assign(data, n);
    s->systemId   = name ? name : "";
    c->priv       = s;
    c->wellFormed = 1;
    c->valid      = 1;
    return c;
}

struct Parser {
    xmlParserCtxtPtr c;
    State&           s;
    explicit Parser(xmlParserCtxtPtr ctxt)
            : c(ctxt), s(*st(ctxt)) {}

    bool eof() const {
        return s.pos >= s.text.size();
    }
    char peek(size_t o = 0) const {
        return s.pos + o < s.text.size() ? s.text[s.pos + o] : '\0';
    }
    void advance(size_t n = 1) {
        for (size_t i = 0; i < n && s.pos < s.text.size(); ++i, ++s.pos) {
            if (s.text[s.pos] == '\n') {
                ++s.line;
                s.col = 1;
            }
            else
                ++s.col;
        }
    }
    bool starts(const char* lit) const {
        return s.text.compare(s.pos, strlen(lit), lit) == 0;
    }
    bool fail(const char* what) {
        c->wellFormed = 0;
        c->errNo      = 1;
        if (c->sax && c->sax->fatalError)
            c->sax->fatalError(c->userData, "%s", what);
        return false;
    }
    static bool isSpace(char ch) {
        return ch == ' ' || ch == '\t' || ch == '\n' || ch == '\r';
    }
    static bool isNameChar(char ch) {
        return isalnum((unsigned char)ch) || ch == '_' || ch == '-' || ch == '.' || ch == ':' || (ch & 0x80);
    }
    void skipSpace() {
        while (!eof() && isSpace(peek()))
            advance();
    }
    std::string name() {
        std::string r;
        while (!eof() && isNameChar(peek())) {
            r += peek();
            advance();
        }
        return r;
    }
    // appends the expansion of the reference at pos (which is at '&') to out
    bool reference(std::string& out) {
        size_t end = s.text.find(';', s.pos);
        if (end == std::string::npos || end - s.pos > 12)
            return fail("unterminated entity reference");
        std::string ent = s.text.substr(s.pos + 1, end - s.pos - 1);
        advance(end - s.pos + 1);
        if (ent == "amp")
            out += '&';
        else if (ent == "lt")
            out += '<';
        else if (ent == "gt")
            out += '>';
        else if (ent == "quot")
            out += '"';
        else if (ent == "apos")
            out += '\'';
        else if (!ent.empty() && ent[0] == '#') {
            unsigned long cp = ent.size() > 1 && (ent[1] == 'x' || ent[1] == 'X') ? strtoul(ent.c_str() + 2, 0, 16)
                                                                                   : strtoul(ent.c_str() + 1, 0, 10);
            if (cp < 0x80)
                out += (char)cp;
            else if (cp < 0x800) {
                out += (char)(0xC0 | (cp >> 6));
                out += (char)(0x80 | (cp & 0x3F));
            }
            else if (cp < 0x10000) {
                out += (char)(0xE0 | (cp >> 12));
                out += (char)(0x80 | ((cp >> 6) & 0x3F));
                out += (char)(0x80 | (cp & 0x3F));
            }
            else {
                out += (char)(0xF0 | (cp >> 18));
                out += (char)(0x80 | ((cp >> 12) & 0x3F));
                out += (char)(0x80 | ((cp >> 6) & 0x3F));
                out += (char)(0x80 | (cp & 0x3F));
            }
        }
        else
            return fail("unknown entity");
        return true;
    }
    void flushText(std::string& t, bool insideRoot) {
        if (t.empty())
            return;
        if (insideRoot && c->sax && c->sax->characters)
            c->sax->characters(c->userData, reinterpret_cast<const xmlChar*>(t.data()), (int)t.size());
        t.clear();
    }

    bool document() {
        xmlSAXHandler* h = c->sax;
        void*          u = c->userData;
        if (h && h->startDocument)
            h->startDocument(u);
        std::vector<std::string> open;
        std::string              text;
        bool                     sawRoot = false;
        while (!eof()) {
            if (peek() != '<') {
                if (peek() == '&') {
                    if (!reference(text))
                        return false;
                }
                else {
                    text += peek();
                    advance();
                }
                continue;
            }
            flushText(text, !open.empty());
            if (starts("<!--")) {
                size_t end = s.text.find("-->", s.pos + 4);
                if (end == std::string::npos)
                    return fail("unterminated comment");
                std::string body = s.text.substr(s.pos + 4, end - s.pos - 4);
                advance(end + 3 - s.pos);
                if (h && h->comment)
                    h->comment(u, reinterpret_cast<const xmlChar*>(body.c_str()));
            }
            else if (starts("<![CDATA[")) {
                size_t end = s.text.find("]]>", s.pos + 9);
                if (end == std::string::npos)
                    return fail("unterminated CDATA section");
                std::string body = s.text.substr(s.pos + 9, end - s.pos - 9);
                advance(end + 3 - s.pos);
                if (h && h->cdataBlock)
                    h->cdataBlock(u, reinterpret_cast<const xmlChar*>(body.data()), (int)body.size());
            }
            else if (starts("<?")) {
                size_t end = s.text.find("?>", s.pos + 2);
                if (end == std::string::npos)
                    return fail("unterminated processing instruction");
                std::string body = s.text.substr(s.pos + 2, end - s.pos - 2);
                advance(end + 2 - s.pos);
                size_t      sp     = body.find_first_of(" \t\r\n");
                std::string target = body.substr(0, sp), data = sp == std::string::npos ? "" : body.substr(sp + 1);
                if (target != "xml" && h && h->processingInstruction)
                    h->processingInstruction(u, reinterpret_cast<const xmlChar*>(target.c_str()),
                                             reinterpret_cast<const xmlChar*>(data.c_str()));
            }
            else if (starts("<!")) {  // DOCTYPE and friends: skipped (with one level of [...] nesting)
                int depth = 0;
                while (!eof()) {
                    char ch = peek();
                    advance();
                    if (ch == '[')
                        ++depth;
                    else if (ch == ']')
                        --depth;
                    else if (ch == '>' && depth <= 0)
                        break;
                }
            }
            else if (starts("</")) {
                advance(2);
                std::string n = name();
                skipSpace();
                if (peek() != '>')
                    return fail("malformed end tag");
                advance();
                if (open.empty() || open.back() != n)
                    return fail("end tag does not match the open element");
                open.pop_back();
                if (h && h->endElement)
                    h->endElement(u, reinterpret_cast<const xmlChar*>(n.c_str()));
            }
            else {
                advance();
                std::string n = name();
                if (n.empty())
                    return fail("malformed start tag");
                if (open.empty() && sawRoot)
                    return fail("more than one root element");
                sawRoot = true;
                std::vector<std::string> kv;
                bool                     empty = false;
                for (;;) {
                    skipSpace();
                    if (eof())
                        return fail("unterminated start tag");
                    if (peek() == '>') {
                        advance();
                        break;
                    }
                    if (peek() == '/' && peek(1) == '>') {
                        advance(2);
                        empty = true;
                        break;
                    }
                    std::string k = name();
                    if (k.empty())
                        return fail("malformed attribute");
                    skipSpace();
                    if (peek() != '=')
                        return fail("attribute without a value");
                    advance();
                    skipSpace();
                    const char q = peek();
                    if (q != '"' && q != '\'')
                        return fail("attribute value is not quoted");
                    advance();
                    std::string v;
                    while (!eof() && peek() != q) {
                        if (peek() == '&') {
                            if (!reference(v))
                                return false;
                        }
                        else {
                            char ch = peek();
                            v += (ch == '\n' || ch == '\t' || ch == '\r') ? ' ' : ch;
                            advance();
                        }
                    }
                    if (eof())
                        return fail("unterminated attribute value");
                    advance();
                    kv.push_back(k);
                    kv.push_back(v);
                }
                std::vector<const xmlChar*> atts;
                for (auto& x : kv)
                    atts.push_back(reinterpret_cast<const xmlChar*>(x.c_str()));
                atts.push_back(0);
                if (h && h->startElement)
                    h->startElement(u, reinterpret_cast<const xmlChar*>(n.c_str()), kv.empty() ? 0 : atts.data());
                if (empty) {
                    if (h && h->endElement)
                        h->endElement(u, reinterpret_cast<const xmlChar*>(n.c_str()));
                }
                else
                    open.push_back(n);
            }
        }
        if (!open.empty())
            return fail("document ends inside an element");
        if (!sawRoot)
            return fail("document is empty");
        if (h && h->endDocument)
            h->endDocument(u);
        return true;
    }
};
}  // namespace

extern "C" {

xmlParserCtxtPtr xmlCreateMemoryParserCtxt(const char* buffer, int size) {
    if (!buffer || size < 0)
        return 0;
    return make_ctxt(buffer, (size_t)size, "<memory>");
}

xmlParserCtxtPtr xmlCreateFileParserCtxt(const char* filename) {
    FILE* f = filename ? fopen(filename, "rb") : 0;
    if (!f)
        return 0;
    std::string data;
    char        buf[65536];
    size_t      n;
    while ((n = fread(buf, 1, sizeof(buf), f)) > 0)
        data.append(buf, n);
    fclose(f);
    return make_ctxt(data.data(), data.size(), filename);
}

xmlParserCtxtPtr xmlCreatePushParserCtxt(xmlSAXHandler* sax, void* userData, const char* chunk, int size,
                                         const char* filename) {
    xmlParserCtxtPtr c = make_ctxt(chunk, chunk && size > 0 ? (size_t)size : 0, filename);
    c->sax             = sax;
    c->userData        = userData;
    return c;
}

int xmlParseChunk(xmlParserCtxtPtr ctxt, const char* chunk, int size, int terminate) {
    if (!ctxt)
        return -1;
    if (chunk && size > 0)
        st(ctxt)->text.append(chunk, (size_t)size);
    if (terminate)
        return xmlParseDocument(ctxt);
    return 0;
}

int xmlParseDocument(xmlParserCtxtPtr ctxt) {
    if (!ctxt)
        return -1;
    Parser p(ctxt);
    return p.document() ? 0 : -1;
}

void xmlFreeParserCtxt(xmlParserCtxtPtr ctxt) {
    if (!ctxt)
        return;
    delete st(ctxt);
    free(ctxt);
}

xmlEntityPtr xmlGetPredefinedEntity(const xmlChar*) {
    return 0; /* predefined entities are expanded by the parser itself */
}

int xmlStrcmp(const xmlChar* a, const xmlChar* b) {
    if (a == b)
        return 0;
    if (!a)
        return -1;
    if (!b)
        return 1;
    return strcmp(reinterpret_cast<const char*>(a), reinterpret_cast<const char*>(b));
}

const xmlChar* xmlSAX2GetSystemId(void* ctx) {
    return reinterpret_cast<const xmlChar*>(st(static_cast<xmlParserCtxtPtr>(ctx))->systemId.c_str());
}
int xmlSAX2GetLineNumber(void* ctx) {
    return st(static_cast<xmlParserCtxtPtr>(ctx))->line;
}
int xmlSAX2GetColumnNumber(void* ctx) {
    return st(static_cast<xmlParserCtxtPtr>(ctx))->col;
}
xmlParserInputPtr resolveEntity(void*, const xmlChar*, const xmlChar*) {
    return 0;
}
xmlDictPtr xmlDictCreate(void) {
    return 0;
}
void xmlDictFree(xmlDictPtr) {}
}
