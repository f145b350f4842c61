/*
 * stsb200.c - the reference's main() sequence (src/sts.c:317-399) for the hot path:
 *     parse_args -> init -> invokeTestSuite -> write_p_val_to_file -> metrics -> destroy      (-m b, -m i)
 *     parse_args -> init -> read_from_p_val_file -> metrics -> destroy                         (-m a -d dir)
 * with the reference's flag letters and meanings for everything the path needs
 * (-i -S -P -T -j -m -d -F -t -w -v -I -O -s; parse_args.c:426-957).  The assessment (result.txt, or the table of
 * finalAnalysisReport.txt with -O) comes from tallies kept on the GPU; the .pvalues files stay compatible with the
 * reference's own sts -m a -d <workDir>.
 */
#define _GNU_SOURCE
#include <getopt.h>
#include <stdlib.h>
#include <string.h>
#include "sts_host.h"

static const char *usage =
	"usage: stsb200 [-v level] [-t test1[,test2]..] [-P num=value[,num=value]..] [-S bitcount] [-i iterations]\n"
	"               [-I reportCycle] [-O] [-s] [-w workDir] [-F r|a] [-j jobnum] [-m b|i|a] [-d pvaluesdir] [-T threads]\n"
	"               [-B batch] [-D device] [-g 0..9] [datafile | -]\n"
	"  same meanings as sts (arcetri/sts 3.2.7); -B (streams per GPU batch) and -D (CUDA device) are new.\n"
	"  -m i iterates on the GPU and writes workDir/sts.JJJJ.<iterations>.<n>.pvalues; -m b iterates and writes the\n"
	"  reference's result.txt (with -O the table of its legacy finalAnalysisReport.txt) from tallies kept on the GPU,\n"
	"  without keeping the p-values; -g 1..9 tests the output of generator g of tools/generators instead of a file:\n"
	"  1, 2, 4, 5 are made on the GPU (any job jumps to its streams), 3, 6, 7, 8, 9 are one chain and run on a host core;\n"
	"  -m a -d dir assesses the .pvalues files of earlier -m i runs, streamed through the GPU tallies in bounded memory.\n";

int
main(int argc, char **argv)
{
	struct sts_state st;
	sts_defaultstate(&st);
	int opt;
	char *brkt, *phrase;
	while ((opt = getopt(argc, argv, "v:t:P:S:i:I:w:F:j:m:T:B:D:d:g:Osh")) != -1) {
		switch (opt) {
		case 'v': st.debuglevel = atoi(optarg); break;
		case 't':
			st.testVectorFlag = true;
			for (phrase = strtok_r(optarg, ",", &brkt); phrase; phrase = strtok_r(NULL, ",", &brkt)) {
				long t = atol(phrase);
				if (t == 0)
					for (int i = 1; i <= STS_NUMOFTESTS; i++) st.testVector[i] = true;
				else if (t < 0 || t > STS_NUMOFTESTS)
					sts_err(1, __func__, "-t test: %ld must be in the range [0-%d]", t, STS_NUMOFTESTS);
				else
					st.testVector[t] = true;
			}
			break;
		case 'P':
			for (phrase = strtok_r(optarg, ",", &brkt); phrase; phrase = strtok_r(NULL, ",", &brkt)) {
				long num, value;
				double d;
				if (sscanf(phrase, "%ld=", &num) != 1 || num < 1 || num > 11)
					sts_err(1, __func__, "-P num=value[,num=value].. num must be in the range [1-11]: %s", phrase);
				if (num <= 9) {
					if (sscanf(phrase, "%ld=%ld", &num, &value) != 2)
						sts_err(1, __func__, "-P expected integer=integer: %s", phrase);
					sts_change_params(&st, num, value, 0.0);
				} else {
					if (sscanf(phrase, "%ld=%lf", &num, &d) != 2)
						sts_err(1, __func__, "-P expected integer=float: %s", phrase);
					sts_change_params(&st, num, 0, d);
				}
			}
			break;
		case 'S': st.tp.n = atol(optarg); break;
		case 'i': st.tp.numOfBitStreams = atol(optarg); break;
		case 'I': st.reportCycle = atol(optarg); break;
		case 'w': free(st.workDir); st.workDir = strdup(optarg); break;
		case 'F':
			if (optarg[0] != 'r' && optarg[0] != 'a')
				sts_err(1, __func__, "-F format: %s must be r or a", optarg);
			st.dataFormat = (enum sts_format) optarg[0];
			break;
		case 'j': st.jobnum = atol(optarg); break;
		case 'm':
			if (optarg[0] == 'b') st.runMode = STS_MODE_ITERATE_AND_ASSESS;
			else if (optarg[0] == 'i') st.runMode = STS_MODE_ITERATE_ONLY;
			else if (optarg[0] == 'a') st.runMode = STS_MODE_ASSESS_ONLY;
			else sts_err(1, __func__, "-m mode: %s must be b, i or a", optarg);
			break;
		case 'd': free(st.pvalues_dir); st.pvalues_dir = strdup(optarg); break;
		case 'g':	/* sts 3 only accepts -g 0 (parse_args.c:480-483); here 1..9 are the generators of tools/generators.c */
			st.generator = atol(optarg);
			if (st.generator < 0 || st.generator > 9)
				sts_err(1, __func__, "-g generator: %ld must be 0 (data file) or 1..9 (tools/generators)", st.generator);
			break;
		case 'T': st.numberOfThreads = atol(optarg); break;
		case 'B': st.batchStreams = atol(optarg); break;
		case 'D': st.device = atoi(optarg); break;
		case 'O': st.legacy_output = true; break;
		case 's': st.resultstxtFlag = true; break;
		case 'h':
		default:
			fputs(usage, stderr);
			return opt == 'h' ? 0 : 1;
		}
	}
	if (optind == argc - 1)
		st.randomDataPath = strdup(argv[optind]);
	else if (optind == argc && st.generator == 1)
		st.randomDataPath = strdup("((tools/generators 1, linear congruential, generated on the GPU))");
	else if (optind == argc && st.generator >= 2) {
		static const char *names[10] = { "", "", "quadratic congruential I", "quadratic congruential II", "cubic congruential", "XOR",
			"modular exponentiation", "Blum-Blum-Shub", "Micali-Schnorr", "SHA-1 based" };
		char label[128];
		snprintf(label, sizeof(label), "((tools/generators %ld, %s, generated on the %s))", st.generator, names[st.generator],
			 (st.generator == 2 || st.generator == 4 || st.generator == 5) ? "GPU" : "host");
		st.randomDataPath = strdup(label);
	}
	else if (optind != argc || st.runMode != STS_MODE_ASSESS_ONLY) {	/* parse_args.c:706-732: -m a needs no data file */
		fputs(usage, stderr);
		return 1;
	}
	if (!st.testVectorFlag)				/* parse_args.c: no -t means all tests */
		for (int i = 1; i <= STS_NUMOFTESTS; i++) st.testVector[i] = true;
	if (st.tp.numOfBitStreams < 1)
		sts_err(1, __func__, "-i iterations: %ld must be > 0", st.tp.numOfBitStreams);
	st.binsIterations = st.tp.numOfBitStreams;
	if ((st.tp.n % 8) != 0)				/* parse_args.c:939-947 */
		sts_err(1, __func__, "bitcount(n): %ld must be a multiple of 8. The added complexity of supporting "
			"a sequence that starts or ends on a non-byte boundary outweighs the "
			"convenience of permitting arbitrary bit lengths", st.tp.n);
	if (st.tp.n < 1000)				/* GLOBAL_MIN_BITCOUNT, defs.h:200 */
		sts_err(1, __func__, "bitcount(n): %ld must >= %d", st.tp.n, 1000);

	if (st.runMode == STS_MODE_ASSESS_ONLY) {	/* sts.c:352-383: init, read the p-value files, metrics */
		if (st.pvalues_dir == NULL)
			sts_err(1, __func__, "-m a requires -d pvaluesdir");
		sts_find_p_val_files(&st);
		sts_init(&st);
		sts_read_from_p_val_file(&st);
		sts_metrics(&st);
		if (st.debuglevel > 0)
			printf("assessed %ld iterations from %s/sts.*.*.%ld.pvalues\n", st.tp.numOfBitStreams, st.pvalues_dir, st.tp.n);
		sts_destroy(&st);
		return 0;
	}
	sts_init(&st);
	sts_invokeTestSuite(&st);
	sts_print(&st);				/* -s */
	if (st.runMode == STS_MODE_ITERATE_ONLY)	/* sts.c:358-360 */
		sts_write_p_val_to_file(&st);
	sts_metrics(&st);			/* -m b: result.txt (or, with -O, the legacy table), tallied on the GPU */
	if (st.debuglevel > 0) {
		for (int t = 1; t <= STS_NUMOFTESTS; t++)
			if (st.testVector[t])
				printf("%-26s count %8ld valid %8ld success %8ld failure %8ld\n", sts_testNames[t], st.count[t],
				       st.valid[t], st.success[t], st.failure[t]);
		if (st.runMode == STS_MODE_ITERATE_ONLY)
			printf("wrote %s/sts.%04ld.%ld.%ld.pvalues\n", st.workDir, st.jobnum, st.tp.numOfBitStreams, st.tp.n);
	}
	sts_destroy(&st);
	return 0;
}
