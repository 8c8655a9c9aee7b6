// fastx.hpp -- minimal streaming FASTA/FASTQ record reader over zlib (plain or gzip'ed files, "-" = stdin).
// Host I/O only (the reference uses kseq.h for this; SURVEY marks it out of the accelerated path).
#pragma once
#include <ctype.h>
#include <string.h>
#include <string>
#include <vector>
#include <zlib.h>

namespace mpb {

struct FastxReader {
	gzFile fp = 0;
	std::vector<unsigned char> buf;
	int64_t pos = 0, fill = 0;
	bool eof = false;
	int last = 0; // first character of the next header, if already consumed
	explicit FastxReader(const char *fn) : buf(1 << 18)
	{
		fp = (fn && strcmp(fn, "-") != 0) ? gzopen(fn, "rb") : gzdopen(0, "rb");
	}
	~FastxReader() { if (fp) gzclose(fp); }
	int getc()
	{
		if (pos >= fill) {
			if (eof) return -1;
			fill = gzread(fp, buf.data(), (unsigned)buf.size());
			pos = 0;
			if (fill <= 0) { eof = true; fill = 0; return -1; }
		}
		return buf[pos++];
	}
	// append the rest of the current line to s (without the line terminator); returns the terminator or -1
	int rest_of_line(std::string &s)
	{
		for (;;) {
			if (pos >= fill) {
				int c = getc();
				if (c < 0) break;
				--pos;
			}
			unsigned char *b = buf.data() + pos, *e = buf.data() + fill;
			unsigned char *nl = (unsigned char*)memchr(b, '\n', (size_t)(e - b));
			if (nl) {
				s.append((const char*)b, (size_t)(nl - b));
				pos = (nl - buf.data()) + 1;
				if (!s.empty() && s.back() == '\r') s.pop_back();
				return '\n';
			}
			s.append((const char*)b, (size_t)(e - b));
			pos = fill;
		}
		return -1;
	}
	// next record; false at end of input
	bool next(std::string &name, std::string &seq)
	{
		int c = last;
		if (c == 0) {
			while ((c = getc()) >= 0 && c != '>' && c != '@') {}
			if (c < 0) return false;
		}
		last = 0;
		name.clear(); seq.clear();
		std::string header;
		rest_of_line(header);
		size_t sp = 0;
		while (sp < header.size() && !isspace((unsigned char)header[sp])) ++sp;
		name.assign(header, 0, sp);
		while ((c = getc()) >= 0 && c != '>' && c != '+' && c != '@') {
			if (c == '\n') continue;
			seq.push_back((char)c);
			rest_of_line(seq);
		}
		if (c == '>' || c == '@') last = c;
		if (c == '+') { // FASTQ: skip the separator line and as many quality characters as there are bases
			std::string junk;
			rest_of_line(junk);
			size_t got = 0;
			while (got < seq.size()) {
				junk.clear();
				if (rest_of_line(junk) < 0 && junk.empty()) break;
				got += junk.size();
			}
		}
		return true;
	}
};


} // namespace mpb
