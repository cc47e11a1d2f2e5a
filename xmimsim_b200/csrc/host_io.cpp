// host_io.cpp -- the reference's on-disk formats without libxml2: XMSI reader, XMSI / XMSO writers, SPE and CSV
// spectrum files.  Replaces xmi_input_read_from_xml_file / xmi_input_write_to_xml_file /
// xmi_output_write_to_xml_file (include/xmi_xml.h; src/xmi_xml.c:966-1263, 1405-1450, 1453-1700, 2553-2580) and the
// history bookkeeping of xmi_output_new (src/xmi_data_structs.c:1369-1519).  Host only.
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <ctime>
#include <map>
#include <memory>
#include <string>
#include <vector>
#include <unistd.h>
#include "engine.h"
#include "xmb_lines.h"

namespace {

// ---- a small XML reader: elements, attributes, character data; comments, PIs and DOCTYPE are skipped ------------
struct Node {
	std::string name, text;
	std::vector<std::pair<std::string, std::string>> attrs;
	std::vector<std::unique_ptr<Node>> kids;
	const Node *child(const char *n) const {
		for (auto &k : kids) if (k->name == n) return k.get();
		return nullptr;
	}
	std::vector<const Node *> children(const char *n) const {
		std::vector<const Node *> v;
		for (auto &k : kids) if (k->name == n) v.push_back(k.get());
		return v;
	}
	const char *attr(const char *n) const {
		for (auto &a : attrs) if (a.first == n) return a.second.c_str();
		return nullptr;
	}
};

std::string unescape(const std::string &s) {
	std::string o;
	for (size_t i = 0; i < s.size(); i++) {
		if (s[i] != '&') { o += s[i]; continue; }
		const size_t e = s.find(';', i);
		if (e == std::string::npos) { o += s[i]; continue; }
		const std::string ent = s.substr(i + 1, e - i - 1);
		if (ent == "lt") o += '<'; else if (ent == "gt") o += '>'; else if (ent == "amp") o += '&';
		else if (ent == "quot") o += '"'; else if (ent == "apos") o += '\'';
		else if (!ent.empty() && ent[0] == '#') o += (char)strtol(ent.c_str() + (ent[1] == 'x' ? 2 : 1), nullptr, ent[1] == 'x' ? 16 : 10);
		else { o += s.substr(i, e - i + 1); }
		i = e;
	}
	return o;
}
std::string escape(const char *s) {
	std::string o;
	for (; s && *s; s++) {
		if (*s == '<') o += "&lt;"; else if (*s == '>') o += "&gt;"; else if (*s == '&') o += "&amp;"; else o += *s;
	}
	return o;
}

// the version string the reference writes into <general version=...> and <xmimsim-results version=...> (VERSION of configure.ac:16; src/xmi_xml.c:1465,1687)
#define XMB_REFERENCE_VERSION "8.1"

struct Parser {
	const std::string &s;
	size_t i = 0;
	int depth = 0;
	std::string err;
	explicit Parser(const std::string &src) : s(src) {}
	bool starts(const char *t) const { return s.compare(i, strlen(t), t) == 0; }
	void skip_misc() {
		for (;;) {
			while (i < s.size() && isspace((unsigned char)s[i])) i++;
			if (starts("<?")) { const size_t e = s.find("?>", i); i = e == std::string::npos ? s.size() : e + 2; }
			else if (starts("<!--")) { const size_t e = s.find("-->", i); i = e == std::string::npos ? s.size() : e + 3; }
			else if (starts("<!DOCTYPE")) {
				int depth = 0;
				for (; i < s.size(); i++) { if (s[i] == '[') depth++; else if (s[i] == ']') depth--; else if (s[i] == '>' && depth <= 0) { i++; break; } }
			} else return;
		}
	}
	std::unique_ptr<Node> element() {
		if (i >= s.size() || s[i] != '<') { err = "expected '<'"; return nullptr; }
		i++;
		std::unique_ptr<Node> n(new Node());
		while (i < s.size() && !isspace((unsigned char)s[i]) && s[i] != '>' && s[i] != '/') n->name += s[i++];
		for (;;) {
			while (i < s.size() && isspace((unsigned char)s[i])) i++;
			if (i >= s.size()) { err = "unterminated tag"; return nullptr; }
			if (s[i] == '/') { i += 2; return n; }
			if (s[i] == '>') { i++; break; }
			std::string an, av;
			while (i < s.size() && s[i] != '=' && !isspace((unsigned char)s[i])) an += s[i++];
			while (i < s.size() && (isspace((unsigned char)s[i]) || s[i] == '=')) i++;
			if (i >= s.size() || (s[i] != '"' && s[i] != '\'')) { err = "attribute " + an + " of " + n->name + " without a quoted value"; return nullptr; }
			const char q = s[i++];
			while (i < s.size() && s[i] != q) av += s[i++];
			if (i >= s.size()) { err = "unterminated attribute value in " + n->name; return nullptr; }
			i++;
			n->attrs.emplace_back(an, unescape(av));
		}
		for (;;) {
			if (i >= s.size()) { err = "unterminated element " + n->name; return nullptr; }
			if (starts("<!--")) { const size_t e = s.find("-->", i); i = e == std::string::npos ? s.size() : e + 3; continue; }
			if (starts("<![CDATA[")) {
				const size_t e = s.find("]]>", i);
				if (e == std::string::npos) { err = "unterminated CDATA section in " + n->name; return nullptr; }
				n->text += s.substr(i + 9, e - i - 9); i = e + 3; continue;
			}
			if (starts("</")) {
				const size_t e = s.find('>', i);
				if (e == std::string::npos) { err = "unterminated end tag of " + n->name; return nullptr; }
				i = e + 1; break;
			}
			if (s[i] == '<') {
				if (++depth > 64) { err = "elements nested deeper than 64 levels"; return nullptr; }
				auto k = element();
				depth--;
				if (!k) return nullptr;
				n->kids.push_back(std::move(k)); continue;
			}
			const size_t e = s.find('<', i);
			n->text += unescape(s.substr(i, e - i));
			i = e;
		}
		return n;
	}
};

std::string trim(const std::string &s) {
	size_t a = 0, b = s.size();
	while (a < b && isspace((unsigned char)s[a])) a++;
	while (b > a && isspace((unsigned char)s[b - 1])) b--;
	return s.substr(a, b - a);
}

struct ReadError { std::string msg; };
double num(const Node *p, const char *tag, const double *dflt = nullptr) {
	const Node *c = p ? p->child(tag) : nullptr;
	if (!c) { if (dflt) return *dflt; throw ReadError{std::string("missing element <") + tag + ">"}; }
	char *end = nullptr;
	const std::string t = trim(c->text);
	const double v = strtod(t.c_str(), &end);
	if (end == t.c_str()) throw ReadError{std::string("bad number in <") + tag + ">"};
	return v;
}
void vec3(const Node *p, const char *tag, double *o) {
	const Node *c = p->child(tag);
	if (!c) throw ReadError{std::string("missing element <") + tag + ">"};
	o[0] = num(c, "x"); o[1] = num(c, "y"); o[2] = num(c, "z");
}
// one <layer>: weight fractions in any positive scale, |w| < 1e-20 dropped, sorted by Z, normalised (src/xmi_xml.c:1209-1263)
void read_layer(const Node *l, xmb_layer &o) {
	std::vector<std::pair<int, double>> el;
	for (const Node *e : l->children("element")) {
		const double w = num(e, "weight_fraction");
		if (std::fabs(w) < 1e-20) continue;
		el.emplace_back((int)num(e, "atomic_number"), w);
	}
	if (el.empty()) throw ReadError{"layer without elements"};
	std::sort(el.begin(), el.end(), [](const std::pair<int, double> &a, const std::pair<int, double> &b) { return a.first < b.first; });
	double sum = 0.0;
	for (auto &e : el) sum += e.second;
	o.n_elements = (int)el.size();
	o.Z = (int *)malloc(sizeof(int) * el.size());
	o.weight = (double *)malloc(sizeof(double) * el.size());
	for (size_t i = 0; i < el.size(); i++) { o.Z[i] = el[i].first; o.weight[i] = el[i].second / sum; }
	o.density = num(l, "density");
	o.thickness = num(l, "thickness");
}
xmb_layer *read_layers(const Node *parent, int *n) {
	*n = 0;
	if (!parent) return nullptr;
	const auto ls = parent->children("layer");
	if (ls.empty()) return nullptr;
	xmb_layer *o = (xmb_layer *)calloc(ls.size(), sizeof(xmb_layer));
	for (size_t i = 0; i < ls.size(); i++) read_layer(ls[i], o[i]);
	*n = (int)ls.size();
	return o;
}

void input_from_node(const Node *root, xmb_input *in) {
	const Node *g = root->child("general");
	if (!g) throw ReadError{"missing <general>"};
	in->general = (xmb_general *)calloc(1, sizeof(xmb_general));
	in->general->version = g->attr("version") ? (float)atof(g->attr("version")) : 1.0f;
	in->general->outputfile = strdup(g->child("outputfile") ? trim(g->child("outputfile")->text).c_str() : "");
	in->general->n_photons_interval = (long)num(g, "n_photons_interval");
	in->general->n_photons_line = (long)num(g, "n_photons_line");
	in->general->n_interactions_trajectory = (int)num(g, "n_interactions_trajectory");
	in->general->comments = strdup(g->child("comments") ? g->child("comments")->text.c_str() : "");
	const Node *c = root->child("composition");
	if (!c) throw ReadError{"missing <composition>"};
	in->composition = (xmb_composition *)calloc(1, sizeof(xmb_composition));
	in->composition->layers = read_layers(c, &in->composition->n_layers);
	in->composition->reference_layer = (int)num(c, "reference_layer");
	const Node *ge = root->child("geometry");
	if (!ge) throw ReadError{"missing <geometry>"};
	xmb_geometry *G = in->geometry = (xmb_geometry *)calloc(1, sizeof(xmb_geometry));
	G->d_sample_source = num(ge, "d_sample_source");
	vec3(ge, "n_sample_orientation", G->n_sample_orientation);
	vec3(ge, "p_detector_window", G->p_detector_window);
	vec3(ge, "n_detector_orientation", G->n_detector_orientation);
	G->area_detector = num(ge, "area_detector");
	G->collimator_height = num(ge, "collimator_height");
	G->collimator_diameter = num(ge, "collimator_diameter");
	G->d_source_slit = num(ge, "d_source_slit");
	G->slit_size_x = num(ge->child("slit_size"), "slit_size_x");
	G->slit_size_y = num(ge->child("slit_size"), "slit_size_y");
	const Node *ex = root->child("excitation");
	if (!ex) throw ReadError{"missing <excitation>"};
	const double zero = 0.0;
	std::vector<xmb_energy_discrete> disc;
	for (const Node *n : ex->children("discrete")) {
		xmb_energy_discrete d{};
		d.energy = num(n, "energy"); d.horizontal_intensity = num(n, "horizontal_intensity"); d.vertical_intensity = num(n, "vertical_intensity");
		d.sigma_x = num(n, "sigma_x", &zero); d.sigma_xp = num(n, "sigma_xp", &zero); d.sigma_y = num(n, "sigma_y", &zero); d.sigma_yp = num(n, "sigma_yp", &zero);
		d.distribution_type = XMB_DISCRETE_MONOCHROMATIC;
		if (const Node *sp = n->child("scale_parameter")) {                       // :777-796
			const std::string t = sp->attr("distribution_type") ? sp->attr("distribution_type") : "monochromatic";
			d.distribution_type = t == "gaussian" ? XMB_DISCRETE_GAUSSIAN : t == "lorentzian" ? XMB_DISCRETE_LORENTZIAN : XMB_DISCRETE_MONOCHROMATIC;
			d.scale_parameter = atof(trim(sp->text).c_str());
		}
		disc.push_back(d);
	}
	std::vector<xmb_energy_continuous> cont;
	for (const Node *n : ex->children("continuous")) {
		xmb_energy_continuous d{};
		d.energy = num(n, "energy"); d.horizontal_intensity = num(n, "horizontal_intensity"); d.vertical_intensity = num(n, "vertical_intensity");
		d.sigma_x = num(n, "sigma_x", &zero); d.sigma_xp = num(n, "sigma_xp", &zero); d.sigma_y = num(n, "sigma_y", &zero); d.sigma_yp = num(n, "sigma_yp", &zero);
		cont.push_back(d);
	}
	std::stable_sort(disc.begin(), disc.end(), [](const xmb_energy_discrete &a, const xmb_energy_discrete &b) { return a.energy < b.energy; });   // :966
	std::stable_sort(cont.begin(), cont.end(), [](const xmb_energy_continuous &a, const xmb_energy_continuous &b) { return a.energy < b.energy; });   // :969
	in->excitation = (xmb_excitation *)calloc(1, sizeof(xmb_excitation));
	in->excitation->n_discrete = (int)disc.size();
	in->excitation->n_continuous = (int)cont.size();
	if (!disc.empty()) { in->excitation->discrete = (xmb_energy_discrete *)malloc(sizeof(xmb_energy_discrete) * disc.size()); memcpy(in->excitation->discrete, disc.data(), sizeof(xmb_energy_discrete) * disc.size()); }
	if (!cont.empty()) { in->excitation->continuous = (xmb_energy_continuous *)malloc(sizeof(xmb_energy_continuous) * cont.size()); memcpy(in->excitation->continuous, cont.data(), sizeof(xmb_energy_continuous) * cont.size()); }
	in->absorbers = (xmb_absorbers *)calloc(1, sizeof(xmb_absorbers));
	if (const Node *ab = root->child("absorbers")) {
		in->absorbers->exc_layers = read_layers(ab->child("excitation_path"), &in->absorbers->n_exc_layers);
		in->absorbers->det_layers = read_layers(ab->child("detector_path"), &in->absorbers->n_det_layers);
	}
	const Node *de = root->child("detector");
	if (!de) throw ReadError{"missing <detector>"};
	xmb_detector *D = in->detector = (xmb_detector *)calloc(1, sizeof(xmb_detector));
	const std::string dt = de->child("detector_type") ? trim(de->child("detector_type")->text) : "";
	if (dt == "SiLi") D->detector_type = XMB_DETECTOR_SILI; else if (dt == "Ge") D->detector_type = XMB_DETECTOR_GE;
	else if (dt == "Si_SDD") D->detector_type = XMB_DETECTOR_SI_SDD; else throw ReadError{"unknown detector_type '" + dt + "'"};   // :1073-1082
	D->live_time = num(de, "live_time");
	D->pulse_width = num(de, "pulse_width");
	const double nch_default = 2048.0;
	D->nchannels = (int)num(de, "nchannels", &nch_default);
	D->gain = num(de, "gain"); D->zero = num(de, "zero"); D->fano = num(de, "fano"); D->noise = num(de, "noise");
	D->crystal_layers = read_layers(de->child("crystal"), &D->n_crystal_layers);
}

// ---- writers ---------------------------------------------------------------------------------------------------------
struct Out {
	FILE *f;
	void open(int depth, const char *tag) { fprintf(f, "%*s<%s>\n", depth, "", tag); }
	void close(int depth, const char *tag) { fprintf(f, "%*s</%s>\n", depth, "", tag); }
	void g(int depth, const char *tag, double v) { fprintf(f, "%*s<%s>%g</%s>\n", depth, "", tag, v, tag); }
	void i(int depth, const char *tag, long v) { fprintf(f, "%*s<%s>%li</%s>\n", depth, "", tag, v, tag); }
	void s(int depth, const char *tag, const char *v) {
		if (!v || !*v) fprintf(f, "%*s<%s/>\n", depth, "", tag);
		else fprintf(f, "%*s<%s>%s</%s>\n", depth, "", tag, escape(v).c_str(), tag);
	}
};

void write_layers(Out &o, int d, const xmb_layer *l, int n) {                         // xmi_write_layer_xml_body, :2553-2580
	for (int i = 0; i < n; i++) {
		o.open(d, "layer");
		double sum = 0.0;
		for (int j = 0; j < l[i].n_elements; j++) sum += l[i].weight[j];
		for (int j = 0; j < l[i].n_elements; j++) {
			if (std::fabs(l[i].weight[j]) < 1E-20) continue;
			o.open(d + 1, "element");
			o.i(d + 2, "atomic_number", l[i].Z[j]);
			o.g(d + 2, "weight_fraction", l[i].weight[j] / sum * 100.0);
			o.close(d + 1, "element");
		}
		o.g(d + 1, "density", l[i].density);
		o.g(d + 1, "thickness", l[i].thickness);
		o.close(d, "layer");
	}
}
void write_vec(Out &o, int d, const char *tag, const double *v) {
	o.open(d, tag); o.g(d + 1, "x", v[0]); o.g(d + 1, "y", v[1]); o.g(d + 1, "z", v[2]); o.close(d, tag);
}
void write_input_body(Out &o, int d, const xmb_input *in) {                           // xmi_write_input_xml_body, :1679-1801
	fprintf(o.f, "%*s<general version=\"%s\">\n", d, "", XMB_REFERENCE_VERSION);
	o.s(d + 1, "outputfile", in->general->outputfile);
	o.i(d + 1, "n_photons_interval", in->general->n_photons_interval);
	o.i(d + 1, "n_photons_line", in->general->n_photons_line);
	o.i(d + 1, "n_interactions_trajectory", in->general->n_interactions_trajectory);
	o.s(d + 1, "comments", in->general->comments);
	o.close(d, "general");
	o.open(d, "composition");
	write_layers(o, d + 1, in->composition->layers, in->composition->n_layers);
	o.i(d + 1, "reference_layer", in->composition->reference_layer);
	o.close(d, "composition");
	const xmb_geometry *G = in->geometry;
	o.open(d, "geometry");
	o.g(d + 1, "d_sample_source", G->d_sample_source);
	write_vec(o, d + 1, "n_sample_orientation", G->n_sample_orientation);
	write_vec(o, d + 1, "p_detector_window", G->p_detector_window);
	write_vec(o, d + 1, "n_detector_orientation", G->n_detector_orientation);
	o.g(d + 1, "area_detector", G->area_detector);
	o.g(d + 1, "collimator_height", G->collimator_height);
	o.g(d + 1, "collimator_diameter", G->collimator_diameter);
	o.g(d + 1, "d_source_slit", G->d_source_slit);
	o.open(d + 1, "slit_size"); o.g(d + 2, "slit_size_x", G->slit_size_x); o.g(d + 2, "slit_size_y", G->slit_size_y); o.close(d + 1, "slit_size");
	o.close(d, "geometry");
	o.open(d, "excitation");
	for (int i = 0; i < in->excitation->n_discrete; i++) {
		const xmb_energy_discrete &e = in->excitation->discrete[i];
		o.open(d + 1, "discrete");
		o.g(d + 2, "energy", e.energy); o.g(d + 2, "horizontal_intensity", e.horizontal_intensity); o.g(d + 2, "vertical_intensity", e.vertical_intensity);
		o.g(d + 2, "sigma_x", e.sigma_x); o.g(d + 2, "sigma_xp", e.sigma_xp); o.g(d + 2, "sigma_y", e.sigma_y); o.g(d + 2, "sigma_yp", e.sigma_yp);
		if (e.distribution_type != XMB_DISCRETE_MONOCHROMATIC)                     // written only when needed, for backwards compatibility (:1742-1752)
			fprintf(o.f, "%*s<scale_parameter distribution_type=\"%s\">%g</scale_parameter>\n", d + 2, "",
			        e.distribution_type == XMB_DISCRETE_GAUSSIAN ? "gaussian" : "lorentzian", e.scale_parameter);
		o.close(d + 1, "discrete");
	}
	for (int i = 0; i < in->excitation->n_continuous; i++) {
		const xmb_energy_continuous &e = in->excitation->continuous[i];
		o.open(d + 1, "continuous");
		o.g(d + 2, "energy", e.energy); o.g(d + 2, "horizontal_intensity", e.horizontal_intensity); o.g(d + 2, "vertical_intensity", e.vertical_intensity);
		o.g(d + 2, "sigma_x", e.sigma_x); o.g(d + 2, "sigma_xp", e.sigma_xp); o.g(d + 2, "sigma_y", e.sigma_y); o.g(d + 2, "sigma_yp", e.sigma_yp);
		o.close(d + 1, "continuous");
	}
	o.close(d, "excitation");
	o.open(d, "absorbers");
	if (in->absorbers->n_exc_layers > 0) { o.open(d + 1, "excitation_path"); write_layers(o, d + 2, in->absorbers->exc_layers, in->absorbers->n_exc_layers); o.close(d + 1, "excitation_path"); }
	if (in->absorbers->n_det_layers > 0) { o.open(d + 1, "detector_path"); write_layers(o, d + 2, in->absorbers->det_layers, in->absorbers->n_det_layers); o.close(d + 1, "detector_path"); }
	o.close(d, "absorbers");
	const xmb_detector *D = in->detector;
	o.open(d, "detector");
	o.s(d + 1, "detector_type", D->detector_type == XMB_DETECTOR_SILI ? "SiLi" : D->detector_type == XMB_DETECTOR_GE ? "Ge" : "Si_SDD");
	o.g(d + 1, "live_time", D->live_time); o.g(d + 1, "pulse_width", D->pulse_width); o.i(d + 1, "nchannels", D->nchannels);
	o.g(d + 1, "gain", D->gain); o.g(d + 1, "zero", D->zero); o.g(d + 1, "fano", D->fano); o.g(d + 1, "noise", D->noise);
	o.open(d + 1, "crystal"); write_layers(o, d + 2, D->crystal_layers, D->n_crystal_layers); o.close(d + 1, "crystal");
	o.close(d, "detector");
}
void write_header(FILE *f, const char *root) {                                      // xmi_write_default_comments, :2124-2150
	char host[256] = "unknown", stamp[64] = "";
	gethostname(host, sizeof(host) - 1);
	const time_t t = time(nullptr);
	struct tm tmv;
	localtime_r(&t, &tmv);
	strftime(stamp, sizeof(stamp), "%F %H:%M:%S (%Z)", &tmv);
	const char *user = getenv("USER") ? getenv("USER") : "unknown";
	fprintf(f, "<?xml version=\"1.0\"?>\n<!DOCTYPE %s SYSTEM \"http://www.xmi.UGent.be/xml/xmimsim-1.0.dtd\">\n", root);
	fprintf(f, "<!-- <Creator>%s (%s)</Creator>\n <Timestamp>%s</Timestamp>\n <Hostname>%s</Hostname>-->\n", user, user, stamp, host);
	fprintf(f, "<!--DO NOT MODIFY THIS FILE UNLESS YOU KNOW WHAT YOU ARE DOING!-->\n");
}

const char *const symbols[] = {"", "H", "He", "Li", "Be", "B", "C", "N", "O", "F", "Ne", "Na", "Mg", "Al", "Si", "P", "S", "Cl", "Ar", "K", "Ca", "Sc", "Ti", "V",
	"Cr", "Mn", "Fe", "Co", "Ni", "Cu", "Zn", "Ga", "Ge", "As", "Se", "Br", "Kr", "Rb", "Sr", "Y", "Zr", "Nb", "Mo", "Tc", "Ru", "Rh", "Pd", "Ag", "Cd", "In", "Sn",
	"Sb", "Te", "I", "Xe", "Cs", "Ba", "La", "Ce", "Pr", "Nd", "Pm", "Sm", "Eu", "Gd", "Tb", "Dy", "Ho", "Er", "Tm", "Yb", "Lu", "Hf", "Ta", "W", "Re", "Os", "Ir",
	"Pt", "Au", "Hg", "Tl", "Pb", "Bi", "Po", "At", "Rn", "Fr", "Ra", "Ac", "Th", "Pa", "U", "Np", "Pu", "Am", "Cm", "Bk", "Cf", "Es", "Fm"};

// <brute_force_history> / <variance_reduction_history> from the raw [100][385][n_int] array (src/xmi_data_structs.c:1400-1519):
// elements of the sample in ascending Z, lines 1..383 with a non-zero sum, orders with counts > 0
void write_history(Out &o, const char *tag, const double *hist, const xmb_input *in, const xmb_xrl_provider *xrl) {
	const int n_int = in->general->n_interactions_trajectory;
	std::vector<int> zs;
	for (int i = 0; i < in->composition->n_layers; i++)
		for (int j = 0; j < in->composition->layers[i].n_elements; j++) zs.push_back(in->composition->layers[i].Z[j]);
	std::sort(zs.begin(), zs.end());
	zs.erase(std::unique(zs.begin(), zs.end()), zs.end());
	auto at = [&](int Z, int line, int k) { return hist[((size_t)(Z - 1) * 385 + (line - 1)) * n_int + (k - 1)]; };
	bool any = false;
	char buf[256];
	if (!hist) zs.clear();
	for (int Z : zs) {
		double tot = 0.0;
		for (int l = 1; l <= 383; l++) for (int k = 1; k <= n_int; k++) tot += at(Z, l, k);
		if (tot == 0.0) continue;
		if (!any) { o.open(1, tag); any = true; }
		snprintf(buf, sizeof(buf), "  <fluorescence_line_counts atomic_number=\"%i\" symbol=\"%s\" total_counts=\"%g\">\n", Z, Z <= 100 ? symbols[Z] : "?", tot);
		fputs(buf, o.f);
		for (int l = 1; l <= 383; l++) {
			double lt = 0.0;
			for (int k = 1; k <= n_int; k++) lt += at(Z, l, k);
			if (lt == 0.0) continue;
			fprintf(o.f, "   <fluorescence_line type=\"%s\" energy=\"%g\" total_counts=\"%g\">\n", xmb_line_name[l], xrl->LineEnergy(Z, -l), lt);
			for (int k = 1; k <= n_int; k++)
				if (at(Z, l, k) > 0.0) fprintf(o.f, "    <counts interaction_number=\"%i\">%g</counts>\n", k, at(Z, l, k));
			fputs("   </fluorescence_line>\n", o.f);
		}
		fputs("  </fluorescence_line_counts>\n", o.f);
	}
	if (any) o.close(1, tag);
	else fprintf(o.f, " <%s/>\n", tag);
}

}  // namespace

// Replaces xmi_input_validate (src/xmi_data_structs.c:899-1255): 0 for a usable input, else the OR of the reference's
// XmiInputFlags (include/xmi_data_structs.h:508-515) of the sections that are not.  One predicate per section.
namespace {
bool layers_invalid(const xmb_layer *l, int n) {
	for (int i = 0; i < n; i++) {
		double sum = 0.0;
		for (int j = 0; j < l[i].n_elements; j++) {
			if (l[i].Z[j] < 1 || l[i].Z[j] > 94) return true;
			if (l[i].weight[j] < 0.0 || l[i].weight[j] > 1.0) return true;
			sum += l[i].weight[j];
		}
		if (sum <= 0.0 || l[i].density <= 0.0 || l[i].thickness <= 0.0) return true;
	}
	return false;
}
bool general_invalid(const xmb_general *g) {
	return !g || g->n_photons_interval <= 0 || g->n_photons_line <= 0 || g->n_interactions_trajectory <= 0 || !g->outputfile ||
	       g->outputfile[0] == 0;
}
bool composition_invalid(const xmb_composition *c) {
	if (!c || c->n_layers < 1 || c->reference_layer < 1 || c->reference_layer > c->n_layers) return true;
	return layers_invalid(c->layers, c->n_layers);
}
bool geometry_invalid(const xmb_geometry *g) {
	return !g || g->n_sample_orientation[2] <= 0.0 || g->d_sample_source <= 0.0 || g->area_detector <= 0.0 || g->collimator_height < 0.0 ||
	       g->collimator_diameter < 0.0 || g->d_source_slit <= 0.0 || g->slit_size_x <= 0.0 || g->slit_size_y <= 0.0;
}
bool excitation_invalid(const xmb_excitation *e) {
	if (!e) return true;
	if ((e->n_discrete == 0 && e->n_continuous < 2) || e->n_continuous == 1) return true;
	for (int i = 0; i < e->n_discrete; i++) {
		const xmb_energy_discrete &d = e->discrete[i];
		if (d.energy <= 0.0 || d.horizontal_intensity < 0.0 || d.vertical_intensity < 0.0 || d.vertical_intensity + d.horizontal_intensity <= 0.0 ||
		    d.sigma_x < 0.0 || d.sigma_y < 0.0 || d.distribution_type < 0 || d.distribution_type > 2 ||
		    (d.distribution_type != 0 && d.scale_parameter <= 0.0))
			return true;
	}
	auto lit = [&](int i) { return e->continuous[i].horizontal_intensity + e->continuous[i].vertical_intensity > 0.0 ? 1 : 0; };
	for (int i = 0; i < e->n_continuous; i++) {
		const xmb_energy_continuous &c = e->continuous[i];
		if (c.energy < 0.0 || c.horizontal_intensity < 0.0 || c.vertical_intensity < 0.0 || c.vertical_intensity + c.horizontal_intensity < 0.0 ||
		    c.sigma_x < 0.0 || c.sigma_y < 0.0)
			return true;
	}
	// a continuum needs intensity somewhere near every interior point: no dark point between dark neighbours, no dark
	// pair at either end (two points: not both dark)
	const int n = e->n_continuous;
	if (n == 2 && e->continuous[0].horizontal_intensity + e->continuous[0].vertical_intensity + e->continuous[1].horizontal_intensity +
	                      e->continuous[1].vertical_intensity == 0.0)
		return true;
	for (int i = 1; n > 2 && i < n - 1; i++) {
		const int before = lit(i - 1), here = lit(i), after = lit(i + 1);
		if ((i == 1 && before + here == 0) || (i == n - 2 && here + after == 0) || before + here + after == 0) return true;
	}
	return false;
}
bool absorbers_invalid(const xmb_absorbers *a) {
	return !a || layers_invalid(a->exc_layers, a->n_exc_layers) || layers_invalid(a->det_layers, a->n_det_layers);
}
bool detector_invalid(const xmb_detector *d) {
	if (!d || d->live_time <= 0.0 || d->pulse_width <= 0.0 || d->gain <= 0.0 || d->fano <= 0.0 || d->noise <= 0.0 ||
	    d->n_crystal_layers < 1 || d->nchannels < 10)
		return true;
	return layers_invalid(d->crystal_layers, d->n_crystal_layers);
}
}  // namespace

extern "C" int xmb_input_validate(const xmb_input *in) {
	if (!in) return 63;
	return (general_invalid(in->general) ? XMB_INPUT_GENERAL : 0) | (composition_invalid(in->composition) ? XMB_INPUT_COMPOSITION : 0) |
	       (geometry_invalid(in->geometry) ? XMB_INPUT_GEOMETRY : 0) | (excitation_invalid(in->excitation) ? XMB_INPUT_EXCITATION : 0) |
	       (absorbers_invalid(in->absorbers) ? XMB_INPUT_ABSORBERS : 0) | (detector_invalid(in->detector) ? XMB_INPUT_DETECTOR : 0);
}

static int input_from_string(const std::string &src, const char *what, xmb_input **input) {
	Parser p(src);
	p.skip_misc();
	std::unique_ptr<Node> root = p.element();
	if (!root) { xmb_set_error("%s: XML syntax error: %s", what, p.err.c_str()); return 0; }
	const Node *body = root.get();
	if (root->name == "xmimsim-results") body = root->child("xmimsim-input");   // an .xmso carries its input
	else if (root->name != "xmimsim") { xmb_set_error("%s: root element is <%s>, expected <xmimsim>", what, root->name.c_str()); return 0; }
	if (!body) { xmb_set_error("%s: no <xmimsim-input>", what); return 0; }
	xmb_input *in = (xmb_input *)calloc(1, sizeof(xmb_input));
	try { input_from_node(body, in); }
	catch (const ReadError &e) { xmb_set_error("%s: %s", what, e.msg.c_str()); xmb_input_free(&in); return 0; }
	// the reference's reader rejects what xmi_input_validate rejects (src/xmi_xml.c:1337-1343)
	if (const int flags = xmb_input_validate(in)) {
		xmb_set_error("%s: error validating input data (sections 0x%x)", what, flags);
		xmb_input_free(&in);
		return 0;
	}
	*input = in;
	return 1;
}

// Replaces xmi_input_read_from_xml_file (src/xmi_xml.c:1289-1340).  *input is malloc'ed (xmb_input_free).  Returns 1 / 0.
extern "C" int xmb_input_read_from_xml_file(const char *xmsifile, xmb_input **input) {
	if (!xmsifile || !input) { xmb_set_error("xmb_input_read_from_xml_file: bad arguments"); return 0; }
	FILE *f = fopen(xmsifile, "rb");
	if (!f) { xmb_set_error("could not open %s", xmsifile); return 0; }
	std::string src;
	char buf[65536];
	size_t n;
	while ((n = fread(buf, 1, sizeof(buf), f)) > 0) src.append(buf, n);
	fclose(f);
	return input_from_string(src, xmsifile, input);
}

// Replaces xmi_input_read_from_xml_string / xmi_input_write_to_xml_string (src/xmi_xml.c:1342-1403): the string form is
// what the solid-angle and escape-ratio caches store as the key of an entry.
extern "C" int xmb_input_read_from_xml_string(const char *xmsistring, xmb_input **input) {
	if (!xmsistring || !input) { xmb_set_error("xmb_input_read_from_xml_string: bad arguments"); return 0; }
	return input_from_string(xmsistring, "xml string", input);
}
extern "C" int xmb_input_write_to_xml_string(const xmb_input *input, char **xmlstring) {
	if (!input || !xmlstring) { xmb_set_error("xmb_input_write_to_xml_string: bad arguments"); return 0; }
	char *buf = nullptr;
	size_t len = 0;
	FILE *f = open_memstream(&buf, &len);
	if (!f) { xmb_set_error("open_memstream failed"); return 0; }
	fputs("<?xml version=\"1.0\"?>\n<!DOCTYPE xmimsim SYSTEM \"http://www.xmi.UGent.be/xml/xmimsim-1.0.dtd\">\n<xmimsim>\n", f);
	Out o{f};
	write_input_body(o, 1, input);
	fputs("</xmimsim>\n", f);
	fclose(f);
	*xmlstring = buf;     // malloc'ed by open_memstream: free()
	return 1;
}

extern "C" void xmb_input_free(xmb_input **p) {
	if (!p || !*p) return;
	xmb_input *d = *p;
	auto fl = [](xmb_layer *l, int n) { if (!l) return; for (int i = 0; i < n; i++) { free(l[i].Z); free(l[i].weight); } free(l); };
	if (d->general) { free(d->general->outputfile); free(d->general->comments); free(d->general); }
	if (d->composition) { fl(d->composition->layers, d->composition->n_layers); free(d->composition); }
	free(d->geometry);
	if (d->excitation) { free(d->excitation->discrete); free(d->excitation->continuous); free(d->excitation); }
	if (d->absorbers) { fl(d->absorbers->exc_layers, d->absorbers->n_exc_layers); fl(d->absorbers->det_layers, d->absorbers->n_det_layers); free(d->absorbers); }
	if (d->detector) { fl(d->detector->crystal_layers, d->detector->n_crystal_layers); free(d->detector); }
	free(d);
	*p = nullptr;
}

// Replaces xmi_input_write_to_xml_file (src/xmi_xml.c:1405-1450).
extern "C" int xmb_input_write_to_xml_file(const xmb_input *input, const char *xmsifile) {
	FILE *f = fopen(xmsifile, "w");
	if (!f) { xmb_set_error("could not write to %s", xmsifile); return 0; }
	write_header(f, "xmimsim");
	fputs("<xmimsim>\n", f);
	Out o{f};
	write_input_body(o, 1, input);
	fputs("</xmimsim>\n", f);
	fclose(f);
	return 1;
}

// The <svg_graphs> block of a stand-alone XMSO file (src/xmi_xml.c:1578-1622 and xmi_write_input_xml_svg, :1880-2022): per
// interaction order a "convoluted" and an "unconvoluted" graphic -- a 500 x 250 box, energy ticks every 5 keV, decade ticks of the
// intensity up to the global maximum of the kind, and one point per channel up to the last channel with at least one count, in
// box coordinates (x linear in energy, y in log10 of the counts, both clamped to the box; single precision as the reference's
// e2c / i2c return float).  The reference's XSLT style sheets draw the spectra from it.
namespace {
const int SVG_W = 500, SVG_H = 250;
float svg_e2c(double energy, const double *energies, int n) {   // n = max_channel, as the reference passes it
	float x = (float)(SVG_W * (energy - energies[0]) / (energies[n - 1] - energies[0]));
	if (x < 0) x = 0;
	if (x > SVG_W) x = SVG_W;
	return x;
}
float svg_i2c(double intensity, double max_log, double min_log) {
	const double ic = intensity < 1 ? 1 : intensity;
	float v = (float)(SVG_H) * (log10(ic) - min_log) / (max_log - min_log);
	if (v < 0) v = 0;
	if (v > SVG_H) v = SVG_H;
	return v;
}
void write_svg_graphic(Out &o, const xmb_input *input, const char *name, int interaction, const double *channels, double maximum) {
	FILE *f = o.f;
	const int nch = input->detector->nchannels;
	const double max_log = log10(maximum), min_log = 0.0;   // minimum = 1
	int max_channel = 0;
	for (int i = nch - 1; i >= 0; i--) if (channels[i] >= 1) { max_channel = i; break; }
	std::vector<double> energies(nch);
	for (int i = 0; i < nch; i++) energies[i] = i * input->detector->gain + input->detector->zero;
	o.open(2, "graphic");
	o.open(3, "id");
	o.s(4, "name", name);
	o.i(4, "interaction", interaction);
	o.close(3, "id");
	o.open(3, "rect");
	fputs("    <view/>\n", f);
	o.open(4, "size");
	o.i(5, "width", SVG_W);
	o.i(5, "height", SVG_H);
	o.g(5, "min_energy", energies[0]);
	o.g(5, "max_energy", energies[max_channel > 0 ? max_channel - 1 : 0]);
	o.close(4, "size");
	o.open(4, "x-axis");
	o.s(5, "name", "Energy (keV)");
	for (double energy = 0.0; energy <= energies[max_channel]; energy += 5.0) {
		o.open(5, "index");
		o.g(6, "value", max_channel > 1 ? svg_e2c(energy, energies.data(), max_channel) : 0.0);
		fprintf(f, "      <name>%.1f</name>\n", energy);
		o.close(5, "index");
	}
	o.close(4, "x-axis");
	o.open(4, "y-axis");
	o.s(5, "name", "Intensity (counts)");
	for (double intensity = 1.0; intensity <= maximum; intensity *= 10.0) {
		o.open(5, "index");
		o.g(6, "value", svg_i2c(intensity, max_log, min_log));
		fprintf(f, "      <name>%.0f</name>\n", intensity);
		o.close(5, "index");
	}
	o.close(4, "y-axis");
	o.close(3, "rect");
	o.open(3, "points");
	o.s(4, "color", "blue");
	for (int i = 0; i <= max_channel; i++) {
		o.open(4, "point");
		o.g(5, "x", max_channel > 1 ? svg_e2c(energies[i], energies.data(), max_channel) : 0.0);
		o.g(5, "y", svg_i2c(channels[i], max_log, min_log));
		o.close(4, "point");
	}
	o.close(3, "points");
	o.close(2, "graphic");
}
}  // namespace

// Replaces xmi_output_new + xmi_output_write_to_xml_file (src/xmi_data_structs.c:1369-1519, src/xmi_xml.c:1453-1700).
// channels_unconv: the raw [(n_int+1)][nch] array of xmb_main_msim; channels_conv: the rows of the detector response
// (index 0 may be NULL unless use_zero_interactions); the two histories [100][385][n_int] (either may be NULL).
// With the <svg_graphs> block (with_svg = 1, as xmi_output_write_to_xml_file passes it) unless XMB_XMSO_NO_SVG is set.
extern "C" int xmb_output_write_to_xml_file(const xmb_input *input, const char *inputfile, const char *xmsofile,
                                            const double *channels_unconv, double *const *channels_conv, const double *brute_history,
                                            const double *var_red_history, int use_zero_interactions, const xmb_xrl_provider *xrl) {
	if (!input || !xmsofile || !channels_unconv || !channels_conv) { xmb_set_error("xmb_output_write_to_xml_file: bad arguments"); return 0; }
	if (!xrl) xrl = xmb_xrl_surrogate();
	FILE *f = fopen(xmsofile, "w");
	if (!f) { xmb_set_error("could not write to %s", xmsofile); return 0; }
	const int n_int = input->general->n_interactions_trajectory, nch = input->detector->nchannels, i0 = use_zero_interactions ? 0 : 1;
	write_header(f, "xmimsim-results");
	fputs("<xmimsim-results version=\"" XMB_REFERENCE_VERSION "\">\n", f);
	Out o{f};
	o.s(1, "inputfile", inputfile ? inputfile : "");
	for (int pass = 0; pass < 2; pass++) {
		o.open(1, pass == 0 ? "spectrum_conv" : "spectrum_unconv");
		for (int j = 0; j < nch; j++) {
			fprintf(f, "  <channel>\n   <channelnr>%i</channelnr>\n   <energy>%g</energy>\n", j, input->detector->gain * j + input->detector->zero);
			for (int i = i0; i <= n_int; i++)
				fprintf(f, "   <counts interaction_number=\"%i\">%g</counts>\n", i, pass == 0 ? channels_conv[i][j] : channels_unconv[(size_t)i * nch + j]);
			fputs("  </channel>\n", f);
		}
		o.close(1, pass == 0 ? "spectrum_conv" : "spectrum_unconv");
	}
	write_history(o, "brute_force_history", brute_history, input, xrl);
	write_history(o, "variance_reduction_history", var_red_history, input, xrl);
	o.open(1, "xmimsim-input");
	write_input_body(o, 2, input);
	o.close(1, "xmimsim-input");
	if (!getenv("XMB_XMSO_NO_SVG")) {
		o.open(1, "svg_graphs");
		double gl_conv = 0.0, gl_unconv = 0.0;   // global maxima over the orders written (maxima[0] = 0 as the reference initialises it)
		for (int i = i0; i <= n_int; i++)
			for (int j = 0; j < nch; j++) {
				gl_conv = std::max(gl_conv, channels_conv[i][j]);
				gl_unconv = std::max(gl_unconv, channels_unconv[(size_t)i * nch + j]);
			}
		for (int i = i0; i <= n_int; i++) {
			write_svg_graphic(o, input, "convoluted", i, channels_conv[i], gl_conv);
			write_svg_graphic(o, input, "unconvoluted", i, channels_unconv + (size_t)i * nch, gl_unconv);
		}
		o.close(1, "svg_graphs");
	}
	fputs("</xmimsim-results>\n", f);
	fclose(f);
	return 1;
}

// SPE and CSV spectrum files of bin/xmimsim.c:546-640.  rows[i] (i = first .. n_int) are spectra of nch channels.
extern "C" int xmb_write_spe_file(const char *filename, const xmb_input *input, const double *row) {
	FILE *f = fopen(filename, "w");
	if (!f) { xmb_set_error("Could not write to %s", filename); return 0; }
	const int nch = input->detector->nchannels;
	fprintf(f, "$SPEC_ID:\n\n$MCA_CAL:\n2\n%g %g\n\n$DATA:\n0\t%i\n", input->detector->zero, input->detector->gain, nch - 1);
	for (int j = 0; j < nch; j++) fprintf(f, "%g%s", row[j], (j + 1) % 8 == 0 ? "\n" : "     ");
	fclose(f);
	return 1;
}
extern "C" int xmb_write_csv_file(const char *filename, const xmb_input *input, double *const *rows, int first) {
	FILE *f = fopen(filename, "w");
	if (!f) { xmb_set_error("Could not write to %s", filename); return 0; }
	const int nch = input->detector->nchannels, n_int = input->general->n_interactions_trajectory;
	for (int j = 0; j < nch; j++) {
		fprintf(f, "%i,%g", j, j * input->detector->gain + input->detector->zero);
		for (int i = first; i <= n_int; i++) fprintf(f, ",%g", rows[i][j]);
		fputc('\n', f);
	}
	fclose(f);
	return 1;
}
